// conv.cu -- K x K integer convolution on interleaved RGB (a6: old/parallel_convolution.c:8-24).
//
//   out[r][c][p] = (sum_{kr,kc} in[r+kr][c+kc][p] * tap[kr][kc]) / norm      (C division, no clamp)
//
// The input is pre-padded and top-left anchored exactly like the reference.  Because the three
// channels are interleaved, output byte column j = 3c+p reads input byte columns j+3*kc: the whole
// thing is a one-channel convolution with horizontal dilation 3, and every kernel below works on
// byte columns.
//
// Two device paths, same results:
//   * separable (taps == outer(u, v), detected on the host; box and binomial/Gaussian kernels):
//     a thread walks down a byte column; per input row it forms the horizontal sum
//     h = sum_kc v[kc]*x[j+3kc] once (K MACs) and feeds it to K rotating vertical accumulators
//     (K MACs): 2K MACs per output instead of K*K.
//   * generic: the same column walk with K*K MACs per input row into K rotating accumulators.
// Taps travel as kernel parameters (constant bank, uniform access).  Accumulation is int32 when
// 255*sum|taps| < 2^31, else int64.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace ompimg {

constexpr int CONV_MAX_K = 15;

struct ConvTaps {
    int32_t t[CONV_MAX_K * CONV_MAX_K];   // generic: t[kr*K+kc]; separable: u = t[0..K), v = t[K..2K)
};

struct ConvParams {
    const uint8_t *in;
    void *out;
    size_t in_pitch, out_pitch;
    int rows, cols3;         // output rows, output byte columns (cols*3)
    int band_rows;
    int32_t norm;
    int out_is_u8;
};

template <typename ACC>
__device__ __forceinline__ void conv_emit(const ConvParams &p, int ro, int j, ACC acc) {
    const ACC q = acc / (ACC)p.norm;     // truncation toward zero, like C
    if (p.out_is_u8) reinterpret_cast<uint8_t *>(p.out)[(size_t)ro * p.out_pitch + j] = (uint8_t)q;
    else *reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(p.out) + (size_t)ro * p.out_pitch + (size_t)j * 4) = (int32_t)q;
}

template <int K, bool SEPARABLE, typename ACC>
__global__ void __launch_bounds__(256) conv_column_kernel(const ConvParams p, const ConvTaps taps) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.cols3) return;
    const int r0 = blockIdx.y * p.band_rows, r1 = min(p.rows, r0 + p.band_rows);
    ACC acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0;
    const uint8_t *col = p.in + j;
    // input rows r0 .. r1+K-2; processed K at a time so that accumulator slots are compile-time
    for (int base = r0; base < r1 + K - 1; base += K) {
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const int ri = base + u;
            if (ri < r1 + K - 1) {
                const uint8_t *row = col + (size_t)ri * p.in_pitch;
                int x[K];
#pragma unroll
                for (int kc = 0; kc < K; ++kc) x[kc] = row[3 * kc];
                if (SEPARABLE) {
                    ACC h = 0;
#pragma unroll
                    for (int kc = 0; kc < K; ++kc) h += (ACC)x[kc] * (ACC)taps.t[K + kc];
#pragma unroll
                    for (int kr = 0; kr < K; ++kr) acc[(u - kr + K) % K] += h * (ACC)taps.t[kr];
                } else {
#pragma unroll
                    for (int kr = 0; kr < K; ++kr) {
                        ACC a = acc[(u - kr + K) % K];
#pragma unroll
                        for (int kc = 0; kc < K; ++kc) a += (ACC)x[kc] * (ACC)taps.t[kr * K + kc];
                        acc[(u - kr + K) % K] = a;
                    }
                }
                // output row ri-(K-1) has now seen all of its K input rows
                const int ro = ri - (K - 1);
                if (ro >= r0) conv_emit<ACC>(p, ro, j, acc[(u + 1) % K]);
                acc[(u + 1) % K] = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path: a thread owns FOUR adjacent output byte columns (one 32-bit word) and walks down a band.
// Per input row it loads the M+1 aligned words that cover its window, builds the three byte-shifted
// copies with PRMT, and forms the horizontal sums with dp4a against "dilated" tap words
// (tap v[kc] sits at byte 3*kc, zeros elsewhere), i.e. four multiply-adds per instruction:
//     h[b] = sum_m dp4a(bytes[b+4m .. b+4m+3], T[m])          b = 0..3
// Separable taps (outer(u, v)): one such sum per row, then K multiply-adds into the K rotating
// accumulators; generic taps: one sum per tap row.  Needs |taps| <= 127 (signed bytes), a 4-byte
// aligned input and int32 accumulation; everything else takes conv_column_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int conv_m(int K) { return (3 * (K - 1) + 1 + 3) / 4; }     // dilated taps span 3(K-1)+1 bytes

struct ConvWordTaps {
    int32_t u[CONV_MAX_K];                        // separable: vertical taps
    uint32_t t[9 * 7];                            // dilated tap words: separable T[m]; generic T[kr][m]
};

// d = c + sum_i (unsigned byte i of a) * (signed byte i of b)
__device__ __forceinline__ int32_t dp4a_us(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int DIVMODE>
__device__ __forceinline__ int32_t conv_div(int32_t acc, int32_t norm, uint32_t magic, int shift) {
    if (DIVMODE == 0) return acc >> shift;                                   // taps >= 0, norm = 2^shift
    if (DIVMODE == 1) return (int32_t)(__umulhi((uint32_t)acc, magic) >> shift);   // taps >= 0, norm > 0: exact for acc < 2^31
    return acc / norm;                                                       // anything: C truncation toward zero
}

template <int K, bool SEPARABLE, int DIVMODE>
__global__ void __launch_bounds__(256, K >= 7 ? 3 : 4) conv_word_kernel(const ConvParams p, const ConvWordTaps taps, uint32_t magic, int shift) {
    constexpr int M = conv_m(K);
    const int jw = blockIdx.x * blockDim.x + threadIdx.x;          // output word index in the row
    const int nwords = (p.cols3 + 3) >> 2;
    if (jw >= nwords) return;
    const int r0 = blockIdx.y * p.band_rows, r1 = min(p.rows, r0 + p.band_rows);
    // words past the end of the row only ever meet zero taps, but they must not be read past the
    // allocation: the (few) threads whose window crosses the pitch re-read the row's last word instead
    const int wmax = (int)(p.in_pitch >> 2) - 1;
    int off[M + 1];
#pragma unroll
    for (int m = 0; m <= M; ++m) off[m] = min(jw + m, wmax) - jw;   // == m for every thread but the last ones of a row
    int32_t acc[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[k][b] = 0;
    const bool full = 4 * jw + 3 < p.cols3;
    const bool inner = jw + M <= wmax;                               // warp-uniform except in the last warp of a row

    // the words of the next input row are requested before the current row is processed (one row of
    // software prefetch: the walk is otherwise exposed to the full memory latency once per row)
    const uint32_t *rowp = reinterpret_cast<const uint32_t *>(p.in + (size_t)r0 * p.in_pitch) + jw;
    const size_t pitch_w = p.in_pitch >> 2;
    uint32_t wn[M + 1];
    auto load_row = [&](const uint32_t *q) {
        if (inner) {
#pragma unroll
            for (int m = 0; m <= M; ++m) wn[m] = __ldg(q + m);
        } else {
#pragma unroll
            for (int m = 0; m <= M; ++m) wn[m] = __ldg(q + off[m]);
        }
    };
    load_row(rowp);
    for (int base = r0; base < r1 + K - 1; base += K) {
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const int ri = base + u;
            if (ri < r1 + K - 1) {
                uint32_t w[M + 1];
#pragma unroll
                for (int m = 0; m <= M; ++m) w[m] = wn[m];
                rowp += pitch_w;
                if (ri + 1 < r1 + K - 1) load_row(rowp);
                // byte-shifted windows: sw[b][m] = bytes [4m+b, 4m+b+4) of the thread's window
                uint32_t sw[4][M];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    sw[0][m] = w[m];
                    sw[1][m] = __byte_perm(w[m], w[m + 1], 0x4321);
                    sw[2][m] = __byte_perm(w[m], w[m + 1], 0x5432);
                    sw[3][m] = __byte_perm(w[m], w[m + 1], 0x6543);
                }
                if (SEPARABLE) {
                    int32_t h[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        int32_t a = 0;
#pragma unroll
                        for (int m = 0; m < M; ++m) a = dp4a_us(sw[b][m], taps.t[m], a);
                        h[b] = a;
                    }
#pragma unroll
                    for (int kr = 0; kr < K; ++kr)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[(u - kr + K) % K][b] += h[b] * taps.u[kr];
                } else {
#pragma unroll
                    for (int kr = 0; kr < K; ++kr)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            int32_t a = acc[(u - kr + K) % K][b];
#pragma unroll
                            for (int m = 0; m < M; ++m) a = dp4a_us(sw[b][m], taps.t[kr * M + m], a);
                            acc[(u - kr + K) % K][b] = a;
                        }
                }
                const int ro = ri - (K - 1);                  // this output row has now seen all of its K input rows
                if (ro >= r0) {
                    int32_t q[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) q[b] = conv_div<DIVMODE>(acc[(u + 1) % K][b], p.norm, magic, shift);
                    if (p.out_is_u8) {
                        uint8_t *o = reinterpret_cast<uint8_t *>(p.out) + (size_t)ro * p.out_pitch + 4 * jw;
                        if (full && (((uintptr_t)o) & 3) == 0)
                            *reinterpret_cast<uint32_t *>(o) = (uint32_t)(q[0] & 255) | ((uint32_t)(q[1] & 255) << 8) | ((uint32_t)(q[2] & 255) << 16) | ((uint32_t)q[3] << 24);
                        else
                            for (int b = 0; b < 4; ++b) if (4 * jw + b < p.cols3) o[b] = (uint8_t)q[b];
                    } else {
                        int32_t *o = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(p.out) + (size_t)ro * p.out_pitch) + 4 * jw;
                        if (full && (((uintptr_t)o) & 15) == 0) *reinterpret_cast<int4 *>(o) = make_int4(q[0], q[1], q[2], q[3]);
                        else
                            for (int b = 0; b < 4; ++b) if (4 * jw + b < p.cols3) o[b] = q[b];
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[(u + 1) % K][b] = 0;
            }
        }
    }
}

template <int K, bool SEP>
static void launch_word(const ConvParams &p, const ConvWordTaps &t, int divmode, uint32_t magic, int shift, dim3 grid, cudaStream_t st) {
    count_launch();
    if (divmode == 0) conv_word_kernel<K, SEP, 0><<<grid, 256, 0, st>>>(p, t, magic, shift);
    else if (divmode == 1) conv_word_kernel<K, SEP, 1><<<grid, 256, 0, st>>>(p, t, magic, shift);
    else conv_word_kernel<K, SEP, 2><<<grid, 256, 0, st>>>(p, t, magic, shift);
}

// any K up to CONV_MAX_K without unrolling (odd sizes nobody benchmarks): one thread per output
template <typename ACC>
__global__ void __launch_bounds__(256) conv_direct_kernel(const ConvParams p, const ConvTaps taps, int K) {
    const long long total = (long long)p.rows * p.cols3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / p.cols3), j = (int)(i - (long long)r * p.cols3);
        ACC a = 0;
        for (int kr = 0; kr < K; ++kr) {
            const uint8_t *row = p.in + (size_t)(r + kr) * p.in_pitch + j;
            for (int kc = 0; kc < K; ++kc) a += (ACC)row[3 * kc] * (ACC)taps.t[kr * K + kc];
        }
        conv_emit<ACC>(p, r, j, a);
    }
}

// taps == outer(u, v) with integer u, v ?  (exact test; u normalised by the gcd of the first non-zero row)
static bool factor_rank1(const int32_t *t, int K, int32_t *u, int32_t *v) {
    int r0 = -1;
    for (int r = 0; r < K && r0 < 0; ++r)
        for (int c = 0; c < K; ++c)
            if (t[r * K + c] != 0) { r0 = r; break; }
    if (r0 < 0) { for (int i = 0; i < K; ++i) u[i] = v[i] = 0; return true; }
    long long g = 0;
    for (int c = 0; c < K; ++c) {
        long long a = t[r0 * K + c] < 0 ? -(long long)t[r0 * K + c] : t[r0 * K + c];
        while (a) { long long tmp = g % a; g = a; a = tmp; }
    }
    for (int c = 0; c < K; ++c) v[c] = (int32_t)(t[r0 * K + c] / g);
    int c0 = 0;
    while (v[c0] == 0) ++c0;
    for (int r = 0; r < K; ++r) {
        if (t[r * K + c0] % v[c0] != 0) return false;
        u[r] = t[r * K + c0] / v[c0];
    }
    for (int r = 0; r < K; ++r)
        for (int c = 0; c < K; ++c)
            if ((long long)u[r] * v[c] != t[r * K + c]) return false;
    return true;
}

template <int K, typename ACC>
static void launch_column(const ConvParams &p, const ConvTaps &taps, bool sep, dim3 grid, cudaStream_t st) {
    if (sep) { count_launch(); conv_column_kernel<K, true, ACC><<<grid, 256, 0, st>>>(p, taps); }
    else { count_launch(); conv_column_kernel<K, false, ACC><<<grid, 256, 0, st>>>(p, taps); }
}

cudaError_t launch_conv(const uint8_t *in, size_t in_pitch, int rows, int cols, int K, const int32_t *taps_host, int32_t norm,
                        void *out, size_t out_pitch, int out_is_u8, int sm_count, cudaStream_t st) {
    if (K < 1 || K > CONV_MAX_K || norm == 0) return cudaErrorInvalidValue;
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    ConvParams p{};
    p.in = in; p.out = out; p.in_pitch = in_pitch; p.out_pitch = out_pitch; p.rows = rows; p.cols3 = cols * 3;
    p.norm = norm; p.out_is_u8 = out_is_u8;
    long long sum_abs = 0;
    for (int i = 0; i < K * K; ++i) sum_abs += taps_host[i] < 0 ? -(long long)taps_host[i] : taps_host[i];
    const bool wide = 255LL * sum_abs >= (1LL << 31);

    ConvTaps taps{};
    int32_t u[CONV_MAX_K], v[CONV_MAX_K];
    const bool sep = K > 1 && factor_rank1(taps_host, K, u, v);
    if (sep) { for (int i = 0; i < K; ++i) { taps.t[i] = u[i]; taps.t[K + i] = v[i]; } }
    else for (int i = 0; i < K * K; ++i) taps.t[i] = taps_host[i];

    // ---- rank-1 taps in range: the planar dp4a/dp2a kernel of conv_sep.cu
    if (sep) {
        cudaError_t e = cudaSuccess;
        if (launch_conv_sep(in, in_pitch, rows, cols, K, u, v, norm, out, out_pitch, out_is_u8, sm_count, st, &e)) return e;
    }

    // ---- word-wide dp4a path
    bool bytes_ok = true, nonneg = true;
    for (int i = 0; i < K * K; ++i) { nonneg = nonneg && taps_host[i] >= 0; }
    if (sep) { for (int i = 0; i < K; ++i) bytes_ok = bytes_ok && v[i] >= -127 && v[i] <= 127; }
    else for (int i = 0; i < K * K; ++i) bytes_ok = bytes_ok && taps_host[i] >= -127 && taps_host[i] <= 127;
    if ((K == 3 || K == 5 || K == 7 || K == 9) && !wide && bytes_ok && (((uintptr_t)in | in_pitch) & 3) == 0) {
        const int M = conv_m(K);
        ConvWordTaps wt{};
        auto dilate = [&](const int32_t *row, uint32_t *dst) {          // tap kc -> byte 3*kc of the M tap words
            for (int m = 0; m < M; ++m) dst[m] = 0;
            for (int kc = 0; kc < K; ++kc) dst[(3 * kc) >> 2] |= (uint32_t)(row[kc] & 255) << (8 * ((3 * kc) & 3));
        };
        if (sep) { for (int i = 0; i < K; ++i) wt.u[i] = u[i]; dilate(v, wt.t); }
        else for (int kr = 0; kr < K; ++kr) dilate(taps_host + kr * K, wt.t + kr * M);
        int divmode = 2, shift = 0;
        uint32_t magic = 0;
        if (nonneg && norm > 0) {
            if ((norm & (norm - 1)) == 0) { divmode = 0; while ((1 << shift) < norm) ++shift; }
            else {
                // floor(a / norm) = umulhi(a, magic) >> shift for every 0 <= a < 2^31 (round-up method, 33-bit-free because a < 2^31)
                int l = 0;
                while ((1LL << l) < norm) ++l;
                const unsigned long long m64 = ((1ULL << (31 + l)) + (unsigned long long)norm - 1) / (unsigned long long)norm;
                if (m64 < (1ULL << 32)) { divmode = 1; magic = (uint32_t)m64; shift = l - 1; }
            }
        }
        const int nwords = (p.cols3 + 3) / 4;
        const int gx = (nwords + 255) / 256;
        int bands = (int)max(1LL, min((long long)(rows + 8 * K - 1) / (8 * K), ((long long)sm_count * 16 + gx - 1) / gx));
        p.band_rows = (rows + bands - 1) / bands;
        bands = (rows + p.band_rows - 1) / p.band_rows;
        const dim3 grid(gx, bands);
        switch (K) {
            case 3: sep ? launch_word<3, true>(p, wt, divmode, magic, shift, grid, st) : launch_word<3, false>(p, wt, divmode, magic, shift, grid, st); break;
            case 5: sep ? launch_word<5, true>(p, wt, divmode, magic, shift, grid, st) : launch_word<5, false>(p, wt, divmode, magic, shift, grid, st); break;
            case 7: sep ? launch_word<7, true>(p, wt, divmode, magic, shift, grid, st) : launch_word<7, false>(p, wt, divmode, magic, shift, grid, st); break;
            case 9: sep ? launch_word<9, true>(p, wt, divmode, magic, shift, grid, st) : launch_word<9, false>(p, wt, divmode, magic, shift, grid, st); break;
        }
        return cudaGetLastError();
    }
    const bool unrolled = (K == 3 || K == 5 || K == 7 || K == 9) && !wide;
    if (unrolled) {
        const int gx = (p.cols3 + 255) / 256;
        // enough bands to fill the machine a few times over, but long enough to amortise the K-1 halo rows
        int bands = (int)max(1LL, min((long long)(rows + 8 * K - 1) / (8 * K), ((long long)sm_count * 32 + gx - 1) / gx));
        p.band_rows = (rows + bands - 1) / bands;
        bands = (rows + p.band_rows - 1) / p.band_rows;
        const dim3 grid(gx, bands);
        switch (K) {
            case 3: launch_column<3, int32_t>(p, taps, sep, grid, st); break;
            case 5: launch_column<5, int32_t>(p, taps, sep, grid, st); break;
            case 7: launch_column<7, int32_t>(p, taps, sep, grid, st); break;
            case 9: launch_column<9, int32_t>(p, taps, sep, grid, st); break;
        }
    } else {
        if (sep) for (int i = 0; i < K * K; ++i) taps.t[i] = taps_host[i];
        const long long total = (long long)rows * p.cols3;
        const int grid = (int)max(1LL, min((total + 255) / 256, (long long)sm_count * 32));
        if (wide) { count_launch(); conv_direct_kernel<long long><<<grid, 256, 0, st>>>(p, taps, K); }
        else { count_launch(); conv_direct_kernel<int32_t><<<grid, 256, 0, st>>>(p, taps, K); }
    }
    return cudaGetLastError();
}

// out = (prior + out) / norm, C division: the accumulate-then-divide of old/parallel_convolution.c:17,20 for callers whose
// output array is not zero on entry (`out` holds the raw tap sums, computed with norm 1)
__global__ void __launch_bounds__(256) accum_div_kernel(int32_t *out, size_t out_pitch, const int32_t *prior, size_t prior_pitch, int rows, int n, int32_t norm) {
    const long long total = (long long)rows * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n, c = i - r * n;
        int32_t *o = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(out) + (size_t)r * out_pitch) + c;
        const int32_t pr = *(reinterpret_cast<const int32_t *>(reinterpret_cast<const uint8_t *>(prior) + (size_t)r * prior_pitch) + c);
        *o = (int32_t)(((long long)pr + (long long)*o) / (long long)norm);
    }
}
cudaError_t launch_accum_div(int32_t *out, size_t out_pitch, const int32_t *prior, size_t prior_pitch, int rows, int n, int32_t norm, int sm_count, cudaStream_t st) {
    if (rows <= 0 || n <= 0) return cudaSuccess;
    const long long total = (long long)rows * n;
    const int grid = (int)std::max(1LL, std::min((total + 255) / 256, (long long)sm_count * 16));
    count_launch();
    accum_div_kernel<<<grid, 256, 0, st>>>(out, out_pitch, prior, prior_pitch, rows, n, norm);
    return cudaGetLastError();
}

}  // namespace ompimg
