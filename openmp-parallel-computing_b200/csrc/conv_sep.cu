// conv_sep.cu -- K x K convolution with rank-1 taps (box, binomial/Gaussian, ...) on interleaved
// uint8 RGB: the fast path of a6 (old/parallel_convolution.c:8-24).  Same results as conv.cu's
// kernels: out[r][c][p] = (sum_{kr,kc} in[r+kr][c+kc][p] * u[kr]*v[kc]) / norm, C division, no clamp,
// pre-padded top-left anchored input.
//
// Data movement: a CTA owns a band of output rows x a strip of NW*128 pixels.  A producer warp
// streams the band's input rows (two per stage) into a shared-memory ring with 1-D bulk asynchronous
// copies (cp.async.bulk -> UBLKCP, mbarrier completion); rows need only 4-byte alignment (the copy
// starts at the 16-byte boundary below the row, the consumers add the row's misalignment).  Each
// consumer lane owns 4 pixels (12 output bytes) and walks down the band; output rows leave through a
// per-warp staging buffer and bulk shared->global copies, so HBM sees full lines in both directions.
//
// Arithmetic per lane and input row (K = 5: 8 input pixels, K = 9: 12):
//   1. de-interleave: 2 PRMT per planar word (r0 r1 r2 r3), (g0..g3), (b0..b3)  -- this removes the
//      dilation-by-3 that interleaved channels impose on the horizontal pass;
//   2. horizontal pass: h[i] = sum_m dp4a(X[m], TH[i][m]) with the taps v pre-shifted into byte
//      position on the host (TH[i][m].byte b = v[4m+b-i]): ceil((K+3)/4) IDP.4A per value, no shifts;
//   3. two consecutive rows' h are packed as u16 pairs (one PRMT); the vertical pass is
//      out[ro] = sum_q dp2a(W[q], UT[q]) over the (K+1)/2 newest pair-rows, kept in a register ring
//      (the row loop is unrolled (K+1)/2 trips so ring slots are compile-time);
//   4. division by the norm: a shift, one multiply-high (exact: host checks the range), a byte
//      select folded into the pack (norm 256 / 65536), or a true division (signed taps).
// K = 5: ~9.5 thread-instructions per output byte against ~33 for conv_word_kernel.
//
// Ranges (checked on the host, anything else takes conv.cu's kernels): unsigned form: taps >= 0,
// v,u <= 255, 255*sum(v) <= 65535 (h fits u16); signed form: |v|,|u| <= 127, 255*sum|v| <= 32767;
// both: 255*sum|u|*sum|v| < 2^31.
#include "common.cuh"
#include "kernels.h"

namespace ompimg {

constexpr int CS_NW = 8;                 // consumer warps per CTA
constexpr int CS_LPX = 4;                // output pixels per lane
constexpr int CS_WPX = 32 * CS_LPX;      // output pixels per warp and row

struct ConvSepParams {
    const uint8_t *in;
    uint8_t *out;
    size_t in_pitch, out_pitch;
    const uint8_t *in_end;       // one past the last input byte
    int rows, cols;
    int band_rows;               // even
    int nw, stages;
    uint32_t row_slot, stage_bytes;
    uint32_t in_row_bytes;       // (cols+K-1)*3
    uint32_t out_bulk;           // output rows are 16-byte aligned: bulk stores
    uint32_t th[4][3];           // horizontal tap words
    uint32_t ut[5];              // vertical tap-pair words: bytes (u[2q], u[2q+1], u[2q-1], u[2q])
    uint32_t magic;
    int32_t shift, norm;
};

__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <bool SG>
__device__ __forceinline__ uint32_t dp4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    if (SG) asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <bool SG, bool HI>
__device__ __forceinline__ uint32_t dp2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    if (SG) {
        if (HI) asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        else asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    } else {
        if (HI) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        else asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    }
    return d;
}

// DIV: 0 acc >> shift | 1 umulhi(acc, magic) | 2 acc / norm (signed) | 3 byte 1 (norm 256, u8 out) | 4 byte 2 (norm 65536, u8 out)
template <int K, bool OUT_U8, int DIV>
__global__ void __launch_bounds__((CS_NW + 1) * 32, K <= 5 ? 3 : 2) conv_sep_kernel(const __grid_constant__ ConvSepParams p) {
    constexpr bool SG = DIV == 2;
    constexpr int NP = (K + 1) / 2;                 // pair-rows in the vertical window
    constexpr int NPX = CS_LPX + K - 1;             // input pixels per lane
    constexpr int NWRD = (NPX * 3 + 3) / 4;         // 32-bit words that hold them
    constexpr int NG = (NPX + 3) / 4;               // planar words per channel
    constexpr int OB = OUT_U8 ? 1 : 4;
    constexpr uint32_t WROW = CS_WPX * 3 * OB;      // staging bytes per warp and output row
    extern __shared__ __align__(128) uint8_t smem[];
    const int NW = p.nw, S = p.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + S * p.stage_bytes;
    const uint32_t empty0 = full0 + 8 * S;
    const uint32_t staging0 = (empty0 + 8 * S + 127u) & ~127u;

    const int band = blockIdx.x, SW = NW * CS_WPX, xs = blockIdx.y * SW;
    const int ya = band * p.band_rows, yb = min(p.rows, ya + p.band_rows);
    const int nin = yb - ya + K - 1;                // input rows ya .. yb+K-2
    const int ntrips = (nin + 1) >> 1;              // two rows per trip (an odd count is padded with a dummy row)

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NW); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ------------------------------------------------------------------ producer warp
        const int rlast = p.rows + K - 2;
        const uint32_t span = min((uint32_t)(SW + K - 1) * 3u, p.in_row_bytes - (uint32_t)xs * 3u);
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < ntrips; ++t) {
            if (t >= S) mbar_wait_relaxed(empty0 + 8 * s, ph ^ 1);
            const uint32_t d0 = smem0 + s * p.stage_bytes;
            const uint8_t *g[2];
            uint32_t nb[2], tx = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = min(ya + 2 * t + j, rlast);
                const uint8_t *gs = p.in + (size_t)r * p.in_pitch + (size_t)xs * 3;
                const uint32_t mis = (uint32_t)(uintptr_t)gs & 15u;
                g[j] = gs - mis;
                nb[j] = (mis + span + 15u) & ~15u;
                if (g[j] + nb[j] > p.in_end) {
                    // the 16-byte granules of this row would end past the image: ordinary byte loads
                    for (uint32_t q = lane; q < span; q += 32) sts8(d0 + j * p.row_slot + mis + q, gs[q]);
                    nb[j] = 0;
                }
                tx += nb[j];
            }
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(full0 + 8 * s, tx);        // also releases the byte stores above
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (nb[j]) bulk_g2s(d0 + j * p.row_slot, g[j], nb[j], full0 + 8 * s);
            }
            if (++s == S) { s = 0; ph ^= 1; }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int seg0 = (xs + warp * CS_WPX) * 3;                                  // first output byte column of the warp
    const int nval = max(0, min(CS_WPX * 3, p.cols * 3 - seg0));                 // valid output byte columns of the warp
    const bool bulk = p.out_bulk && nval == CS_WPX * 3;
    const int lval = nval - lane * 12;                                          // valid output values of the lane (<= 0: none)
    const uint32_t my = keep(smem0 + (uint32_t)(warp * CS_WPX + lane * CS_LPX) * 3u);
    const uint32_t row_slot = keep(p.row_slot);
    const uint32_t stage_bytes = keep(p.stage_bytes);
    const uint32_t ring_bytes = keep((uint32_t)S * p.stage_bytes);
    const uint32_t eoff = keep(8u * (uint32_t)S);
    const uint32_t pitch_lo = keep((uint32_t)p.in_pitch);
    const uint32_t stg = keep(staging0 + (uint32_t)warp * (4u * WROW) + (uint32_t)lane * (12u * OB));
    const uint32_t stg_warp = staging0 + (uint32_t)warp * (4u * WROW);
    uint32_t mrow = (uint32_t)(uintptr_t)p.in + (uint32_t)ya * pitch_lo;        // low address bits of the current input row
    uint8_t *orow = p.out + (size_t)ya * p.out_pitch + (size_t)seg0 * OB;       // next output row of this warp
    int ro = ya;

    uint32_t W[NP][12];
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int c = 0; c < 12; ++c) W[q][c] = 0u;

    uint32_t so = 0, fb = full0, ph = 0, sbuf = 0;
    for (int t0 = 0; t0 < ntrips; t0 += NP) {
#pragma unroll
        for (int s = 0; s < NP; ++s) {
            const int t = t0 + s;
            if (t < ntrips) {
                mbar_wait(fb, ph);
                uint32_t h[2][12];
#pragma unroll
                for (int rho = 0; rho < 2; ++rho) {
                    const uint32_t a = my + so + (uint32_t)rho * row_slot + (mrow & 15u);
                    mrow += pitch_lo;
                    uint32_t w[NWRD];
#pragma unroll
                    for (int i = 0; i < NWRD; ++i) w[i] = lds32(a + 4 * i);
                    if (rho == 1) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(fb + eoff);     // both rows of the stage are in registers
                    }
                    // planar words: X[ch][m] = channel ch of pixels 4m .. 4m+3 (pixels past NPX meet zero taps)
                    uint32_t X[3][NG];
#pragma unroll
                    for (int m = 0; m < NG; ++m) {
                        const uint32_t wa = w[3 * m];
                        const uint32_t wb = 3 * m + 1 < NWRD ? w[3 * m + 1] : 0u;
                        X[0][m] = prmt(wa, wb, 0x0630);     // r0 r1 r2 .
                        X[1][m] = prmt(wa, wb, 0x0741);     // g0 g1 g2 .
                        X[2][m] = prmt(wa, wb, 0x0052);     // b0 b1 . .
                        if (3 * m + 2 < NWRD) {
                            const uint32_t wc = w[3 * m + 2];
                            X[0][m] = prmt(X[0][m], wc, 0x5210);   // r3 = wc.b1
                            X[1][m] = prmt(X[1][m], wc, 0x6210);   // g3 = wc.b2
                            X[2][m] = prmt(X[2][m], wc, 0x7410);   // b2 = wc.b0, b3 = wc.b3
                        }
                    }
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t acc = 0u;
#pragma unroll
                            for (int m = 0; m < NG; ++m)
                                if (4 * m + 3 - i >= 0 && 4 * m - i < K) acc = dp4<SG>(X[ch][m], p.th[i][m < 3 ? m : 0], acc);
                            h[rho][ch * 4 + i] = acc;
                        }
                }
#pragma unroll
                for (int c = 0; c < 12; ++c) W[s][c] = prmt(h[0][c], h[1][c], 0x5410);

                if (t >= (K - 1) / 2) {
                    // output rows ro (window starts on the even row of a pair) and ro+1 (odd row)
                    if (bulk) {
                        if (lane == 0) bulk_wait_read<1>();          // the staging buffer of two trips ago has left
                        __syncwarp();
                    }
                    const bool two = ro + 1 < yb;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (half == 1 && !two) break;
                        int32_t q[12];
#pragma unroll
                        for (int c = 0; c < 12; ++c) {
                            uint32_t acc = 0u;
#pragma unroll
                            for (int k = 0; k < NP; ++k) {
                                const uint32_t wq = W[(s + 1 + k) % NP][c];
                                acc = half ? dp2<SG, true>(wq, p.ut[k], acc) : dp2<SG, false>(wq, p.ut[k], acc);
                            }
                            if (DIV == 0) q[c] = (int32_t)(acc >> p.shift);
                            else if (DIV == 1) q[c] = (int32_t)__umulhi(acc, p.magic);
                            else if (DIV == 2) q[c] = (int32_t)acc / p.norm;
                            else q[c] = (int32_t)acc;
                        }
                        uint8_t *grow = orow + (size_t)half * p.out_pitch;
                        // value j of the lane (byte column 12*lane + j) is pixel j/3, channel j%3
#define CS_Q(j) q[((j) % 3) * 4 + (j) / 3]
                        if (OUT_U8) {
                            constexpr uint32_t SEL = DIV == 3 ? 0x0051u : DIV == 4 ? 0x0062u : 0x0040u;
                            uint32_t o[3];
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                o[k] = prmt(prmt(CS_Q(4 * k), CS_Q(4 * k + 1), SEL), prmt(CS_Q(4 * k + 2), CS_Q(4 * k + 3), SEL), 0x5410);
                            if (bulk) {
                                const uint32_t d = stg + (sbuf * 2u + half) * WROW;
                                sts32(d, o[0]); sts32(d + 4, o[1]); sts32(d + 8, o[2]);
                            } else if (lval > 0) {
                                uint8_t *gp = grow + lane * 12;
                                if (lval >= 12 && ((uintptr_t)gp & 3) == 0) {
                                    reinterpret_cast<uint32_t *>(gp)[0] = o[0];
                                    reinterpret_cast<uint32_t *>(gp)[1] = o[1];
                                    reinterpret_cast<uint32_t *>(gp)[2] = o[2];
                                } else {
#pragma unroll
                                    for (int j = 0; j < 12; ++j)
                                        if (j < lval) gp[j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
                                }
                            }
                        } else {
                            if (bulk) {
                                const uint32_t d = stg + (sbuf * 2u + half) * WROW;
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    sts128(d + 16 * k, make_uint4((uint32_t)CS_Q(4 * k), (uint32_t)CS_Q(4 * k + 1), (uint32_t)CS_Q(4 * k + 2), (uint32_t)CS_Q(4 * k + 3)));
                            } else if (lval > 0) {
                                int32_t *gp = reinterpret_cast<int32_t *>(grow) + lane * 12;
#pragma unroll
                                for (int j = 0; j < 12; ++j)
                                    if (j < lval) gp[j] = CS_Q(j);
                            }
                        }
#undef CS_Q
                    }
                    if (bulk) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            const uint32_t d = stg_warp + sbuf * 2u * WROW;
                            bulk_s2g(orow, d, WROW);
                            if (two) bulk_s2g(orow + p.out_pitch, d + WROW, WROW);
                            bulk_commit();
                        }
                        sbuf ^= 1u;
                    }
                    orow += 2 * p.out_pitch;
                    ro += 2;
                }
                so += stage_bytes;
                fb += 8;
                if (so == ring_bytes) { so = 0; fb = full0; ph ^= 1; }
            }
        }
    }
    if (bulk && lane == 0) bulk_wait_all<0>();
}

template <int K, bool OUT_U8>
static cudaError_t launch_sep_k(const ConvSepParams &p, int div, dim3 grid, int smem, cudaStream_t st) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        count_launch();
        kern<<<grid, (p.nw + 1) * 32, smem, st>>>(p);
        return cudaGetLastError();
    };
    switch (div) {
        case 0: return go(conv_sep_kernel<K, OUT_U8, 0>);
        case 1: return go(conv_sep_kernel<K, OUT_U8, 1>);
        case 2: return go(conv_sep_kernel<K, OUT_U8, 2>);
        case 3: if (OUT_U8) return go(conv_sep_kernel<K, true, 3>); break;
        case 4: if (OUT_U8) return go(conv_sep_kernel<K, true, 4>); break;
    }
    return cudaErrorInvalidValue;
}

// Returns false when the taps / layout are outside this path's ranges (caller uses conv.cu's kernels).
bool launch_conv_sep(const uint8_t *in, size_t in_pitch, int rows, int cols, int K, const int32_t *u, const int32_t *v, int32_t norm,
                     void *out, size_t out_pitch, int out_is_u8, int sm_count, cudaStream_t st, cudaError_t *err) {
    if (!(K == 3 || K == 5 || K == 7 || K == 9)) return false;
    if ((((uintptr_t)in | in_pitch) & 3) != 0) return false;
    const unsigned long long in_row = (unsigned long long)(cols + K - 1) * 3ull;
    if (in_row > 0x7fffffffull || in_pitch < in_row) return false;
    long long su = 0, sv = 0, mu = 0, mv = 0;
    bool nonneg = norm > 0;
    for (int i = 0; i < K; ++i) {
        const long long au = u[i] < 0 ? -(long long)u[i] : u[i], av = v[i] < 0 ? -(long long)v[i] : v[i];
        su += au; sv += av; mu = au > mu ? au : mu; mv = av > mv ? av : mv;
        nonneg = nonneg && u[i] >= 0 && v[i] >= 0;
    }
    const long long acc_max = 255ll * su * sv;
    if (acc_max >= (1ll << 31)) return false;
    int div = -1, shift = 0;
    uint32_t magic = 0;
    if (nonneg && mu <= 255 && mv <= 255 && 255ll * sv <= 65535) {
        if ((norm & (norm - 1)) == 0) {
            while ((1 << shift) < norm) ++shift;
            div = out_is_u8 && shift == 8 ? 3 : out_is_u8 && shift == 16 ? 4 : 0;
        } else {
            // floor(a / norm) == umulhi(a, m), m = ceil(2^32 / norm), for every a with a * (m*norm - 2^32) < 2^32
            const unsigned long long m = ((1ull << 32) + (unsigned long long)norm - 1) / (unsigned long long)norm;
            const unsigned long long e = m * (unsigned long long)norm - (1ull << 32);
            if (m < (1ull << 32) && (unsigned long long)acc_max * e < (1ull << 32)) { div = 1; magic = (uint32_t)m; }
        }
    }
    if (div < 0) {
        if (mu <= 127 && mv <= 127 && 255ll * sv <= 32767) div = 2;
        else return false;
    }

    ConvSepParams p{};
    p.in = in; p.out = reinterpret_cast<uint8_t *>(out); p.in_pitch = in_pitch; p.out_pitch = out_pitch;
    p.in_end = in + (size_t)(rows + K - 2) * in_pitch + (size_t)in_row;
    p.rows = rows; p.cols = cols;
    p.in_row_bytes = (uint32_t)in_row;
    p.magic = magic; p.shift = shift; p.norm = norm;
    p.out_bulk = (((uintptr_t)out | out_pitch) & 15) == 0;
    for (int i = 0; i < 4; ++i)
        for (int m = 0; m < 3; ++m) {
            uint32_t wv = 0;
            for (int b = 0; b < 4; ++b) {
                const int k = 4 * m + b - i;
                if (k >= 0 && k < K) wv |= (uint32_t)(v[k] & 255) << (8 * b);
            }
            p.th[i][m] = wv;
        }
    auto ub = [&](int k) -> uint32_t { return k >= 0 && k < K ? (uint32_t)(u[k] & 255) : 0u; };
    for (int q = 0; q < 5; ++q) p.ut[q] = ub(2 * q) | ub(2 * q + 1) << 8 | ub(2 * q - 1) << 16 | ub(2 * q) << 24;

    p.nw = min(CS_NW, (cols + CS_WPX - 1) / CS_WPX);
    const int SW = p.nw * CS_WPX;
    const int nstrips = (cols + SW - 1) / SW;
    if (nstrips > 65535) return false;
    p.row_slot = (uint32_t)((16 + (SW + K - 1) * 3 + 16 + 127) & ~127);
    p.stage_bytes = 2 * p.row_slot;
    p.stages = 4;
    const int ob = out_is_u8 ? 1 : 4;
    const int smem = p.stages * (int)p.stage_bytes + 16 * p.stages + 128 + p.nw * 4 * CS_WPX * 3 * ob;
    // bands: an even number of rows, sized so that the CTAs fill a whole number of waves of resident CTAs
    const int occ = K <= 5 ? 3 : 2;
    const long long slots = (long long)sm_count * occ;
    int best_rows = rows, best_score = 1 << 30;
    for (int wv = 1; wv <= 64; ++wv) {
        const long long nb = (wv * slots) / nstrips;
        if (nb < 1) continue;
        int br = (int)((rows + nb - 1) / nb);
        br += br & 1;
        if (br < 2 * K) break;
        const int score = br > 96 ? br - 96 : 96 - br;
        if (score < best_score) { best_score = score; best_rows = br; }
    }
    best_rows += best_rows & 1;
    p.band_rows = best_rows;
    const int nbands = (rows + p.band_rows - 1) / p.band_rows;
    const dim3 grid((unsigned)nbands, (unsigned)nstrips);
    switch (K) {
        case 3: *err = out_is_u8 ? launch_sep_k<3, true>(p, div, grid, smem, st) : launch_sep_k<3, false>(p, div, grid, smem, st); break;
        case 5: *err = out_is_u8 ? launch_sep_k<5, true>(p, div, grid, smem, st) : launch_sep_k<5, false>(p, div, grid, smem, st); break;
        case 7: *err = out_is_u8 ? launch_sep_k<7, true>(p, div, grid, smem, st) : launch_sep_k<7, false>(p, div, grid, smem, st); break;
        default: *err = out_is_u8 ? launch_sep_k<9, true>(p, div, grid, smem, st) : launch_sep_k<9, false>(p, div, grid, smem, st); break;
    }
    return true;
}

}  // namespace ompimg
