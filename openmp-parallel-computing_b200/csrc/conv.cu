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

}  // namespace ompimg
