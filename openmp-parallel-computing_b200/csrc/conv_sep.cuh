// conv_sep.cuh -- K x K convolution with rank-1 taps (box, binomial/Gaussian, ...) on interleaved
// uint8 RGB: the fast path of a6 (old/parallel_convolution.c:8-24).  Same results as conv.cu's
// kernels: out[r][c][p] = (sum_{kr,kc} in[r+kr][c+kc][p] * u[kr]*v[kc]) / norm, C division, no clamp,
// pre-padded top-left anchored input.  One translation unit per K (conv_sep_k*.cu) instantiates it.
//
// Data movement: a CTA owns a band of output rows x a strip of NW*128 pixels.  A producer warp
// streams the band's input rows (two per stage) into a shared-memory ring with 1-D bulk asynchronous
// copies (cp.async.bulk -> UBLKCP, mbarrier completion); rows need only 4-byte alignment (the copy
// starts at the 16-byte boundary below the row, the consumers add the row's misalignment).  Each
// consumer lane owns 4 pixels (12 output bytes) and walks down the band; output rows are transposed
// through a per-warp staging buffer so that the warp stores full 16-byte pieces of contiguous rows.
// The ring has a multiple of (K+1)/2 stages, so the stage of an unrolled trip is a compile-time
// offset (no wrap test per trip).
//
// Arithmetic per lane and input row (K = 5: 8 input pixels, K = 9: 12):
//   1. de-interleave: 2 PRMT per planar word (r0 r1 r2 r3), (g0..g3), (b0..b3)  -- this removes the
//      dilation-by-3 that interleaved channels impose on the horizontal pass;
//   2. horizontal pass: h[i] = sum_m dp4a(X[m], TH[i][m]) with the taps v pre-shifted into byte
//      position on the host (TH[i][m].byte b = v[4m+b-i]): ceil((K+3)/4) IDP.4A per value, no shifts;
//   3. vertical pass, two forms:
//      general u (VBOX = false): two consecutive rows' h are packed as u16 pairs (one PRMT) and
//        out[ro] = sum_q dp2a(W[q], UT[q]) over the (K+1)/2 newest pair-rows, kept in a register ring;
//      constant u (VBOX = true, box blur): two neighbouring columns' h are packed as u16 pairs and a
//        running window sum S += P[r] - P[r-K] (ONE IADD3 per pair and row, both halves at once; the
//        last K+1 rows of P stay in a register ring) replaces the K multiply-adds;
//      the row loop is unrolled (K+1)/2 trips so that ring slots are compile-time registers;
//   4. division by the norm: a shift, one multiply-high (exact: the host checks the range), a byte
//      select folded into the pack (norm 256 / 65536), or a true division (signed taps).
//
// Ranges (checked on the host, anything else takes conv.cu's kernels): unsigned form: taps >= 0,
// v,u <= 255, 255*sum(v) <= 65535 (h fits u16); signed form: |v|,|u| <= 127, 255*sum|v| <= 32767;
// both: 255*sum|u|*sum|v| < 2^31; VBOX: the window sum fits u16 as well.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ompimg {

#ifndef CS_MINB5
#define CS_MINB5 4                     // resident CTAs per SM the K <= 5 kernels are compiled for (K > 5: 3)
#endif
#ifndef CS_NWARPS
#define CS_NWARPS 4
#endif
constexpr int CS_NW = CS_NWARPS;         // consumer warps per CTA
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
    int relaxed;                 // producer sleeps between probes of the empty barrier
    uint32_t row_slot, stage_bytes;
    uint32_t in_row_bytes;       // (cols+K-1)*3
    uint32_t out_wide;           // output rows are 16-byte aligned: 128-bit stores through the staging buffer
    uint32_t th[4][3];           // horizontal tap words
    uint32_t ut[5];              // vertical tap-pair words: bytes (u[2q], u[2q+1], u[2q-1], u[2q])
    uint32_t magic;
    int32_t shift, norm;
};

template <int K>
cudaError_t launch_conv_sep_k(const ConvSepParams &p, int out_u8, int div, int vbox, dim3 grid, int smem, cudaStream_t st);

#ifdef OMPIMG_CONV_SEP_IMPL

__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// one lane of a converged warp (always the same one for the full mask); the compiler then keeps the
// elected lane's address arithmetic in uniform registers, which bulk copies take directly
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
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
// VBOX (DIV 0 or 1 only): see the header comment.
template <int K, bool OUT_U8, int DIV, bool VBOX>
__global__ void __launch_bounds__((CS_NW + 1) * 32, K <= 5 ? CS_MINB5 : 3) conv_sep_kernel(const __grid_constant__ ConvSepParams p) {
    constexpr bool SG = DIV == 2;
    constexpr int NP = (K + 1) / 2;                 // trips per unrolled body; pair-rows in the vertical window
    constexpr int NPX = CS_LPX + K - 1;             // input pixels per lane
    constexpr int NWRD = (NPX * 3 + 3) / 4;         // 32-bit words that hold them
    constexpr int NG = (NPX + 3) / 4;               // planar words per channel
    constexpr int OB = OUT_U8 ? 1 : 4;
    constexpr uint32_t WROW = CS_WPX * 3 * OB;      // staging bytes per warp and output row
    constexpr int RING = VBOX ? (K + 1) * 6 : NP * 12;
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
        const uint8_t *gs = p.in + (size_t)ya * p.in_pitch + (size_t)xs * 3;     // start of the strip in the current row
        int r = ya;
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < ntrips; ++t) {
            if (t >= S) { if (p.relaxed) mbar_wait_relaxed(empty0 + 8 * s, ph ^ 1); else mbar_wait(empty0 + 8 * s, ph ^ 1); }
            const uint32_t d0 = smem0 + s * p.stage_bytes;
            // a dummy row (odd row count) re-sends the row before it
            const uint8_t *g1 = r + 1 <= rlast ? gs + p.in_pitch : gs;
            const uint32_t mis0 = (uint32_t)(uintptr_t)gs & 15u, mis1 = (uint32_t)(uintptr_t)g1 & 15u;
            const uint32_t nb0 = (mis0 + span + 15u) & ~15u, nb1 = (mis1 + span + 15u) & ~15u;
            const bool plain0 = gs - mis0 + nb0 > p.in_end, plain1 = g1 - mis1 + nb1 > p.in_end;
            if (plain0 | plain1) {
                // the 16-byte granules of a row would end past the image (its last row): ordinary byte loads
                if (plain0) for (uint32_t q = lane; q < span; q += 32) sts8(d0 + mis0 + q, gs[q]);
                if (plain1) for (uint32_t q = lane; q < span; q += 32) sts8(d0 + p.row_slot + mis1 + q, g1[q]);
                __syncwarp();
            }
            if (elect_one()) {
                mbar_expect_tx(full0 + 8 * s, (plain0 ? 0u : nb0) + (plain1 ? 0u : nb1));   // also releases the byte stores above
                if (!plain0) bulk_g2s(d0, gs - mis0, nb0, full0 + 8 * s);
                if (!plain1) bulk_g2s(d0 + p.row_slot, g1 - mis1, nb1, full0 + 8 * s);
            }
            gs += 2 * p.in_pitch;
            r += 2;
            if (r > rlast) { gs -= (size_t)(r - rlast) * p.in_pitch; r = rlast; }
            if (++s == S) { s = 0; ph ^= 1; }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int seg0 = (xs + warp * CS_WPX) * 3;                                  // first output byte column of the warp
    const int nval = max(0, min(CS_WPX * 3, p.cols * 3 - seg0));                 // valid output byte columns of the warp
    const bool wide = p.out_wide && nval == CS_WPX * 3;
    const int lval = nval - lane * 12;                                          // valid output values of the lane (<= 0: none)
    const uint32_t my = keep(smem0 + (uint32_t)(warp * CS_WPX + lane * CS_LPX) * 3u);
    const uint32_t row_slot = keep(p.row_slot);
    const uint32_t stage_bytes = keep(p.stage_bytes);
    const uint32_t ring_bytes = (uint32_t)S * p.stage_bytes;
    const uint32_t eoff = 8u * (uint32_t)S;
    const uint32_t pitch_lo = keep((uint32_t)p.in_pitch);
    // staging: two output rows per warp; written as 12 values per lane, read back as 16-byte pieces
    // (piece c = lane + 32*j lies in row c / PPR at byte (c % PPR) * 16; j < NPC)
    constexpr int PPR = (int)WROW / 16, NPC = 2 * PPR / 32 + (2 * PPR % 32 ? 1 : 0);
    const uint32_t stg_warp = staging0 + (uint32_t)warp * (2u * WROW);
    const uint32_t stg = stg_warp + (uint32_t)lane * (12u * OB);
    const uint32_t stg_rd = stg_warp + (uint32_t)lane * 16u;
    uint32_t mrow = (uint32_t)(uintptr_t)p.in + (uint32_t)ya * pitch_lo;        // low address bits of the current input row
    uint8_t *orow = p.out + (size_t)ya * p.out_pitch + (size_t)seg0 * OB;       // next output row of this warp
    int ro = ya;

    uint32_t W[RING];            // VBOX: P[slot][6] (slot = input row mod K+1); else W[pair-row slot][12]
#pragma unroll
    for (int q = 0; q < RING; ++q) W[q] = 0u;
    uint32_t SUM[6];             // VBOX: running window sums, two columns per register
#pragma unroll
    for (int q = 0; q < 6; ++q) SUM[q] = 0u;

    uint32_t so = 0, fb = full0, ph = 0;           // ring position of the unrolled body's first trip
    for (int t0 = 0; t0 < ntrips; t0 += NP) {
#pragma unroll
        for (int s = 0; s < NP; ++s) {
            const int t = t0 + s;
            if (t < ntrips) {
                mbar_wait(fb + 8 * s, ph);
                uint32_t w[2][NWRD];
#pragma unroll
                for (int rho = 0; rho < 2; ++rho) {
                    const uint32_t a = my + so + (uint32_t)s * stage_bytes + (uint32_t)rho * row_slot + (mrow & 15u);
                    mrow += pitch_lo;
#pragma unroll
                    for (int i = 0; i < NWRD; ++i) w[rho][i] = lds32(a + 4 * i);
                }
                __syncwarp();                                  // also: every lane has read the previous trip's staging rows
                if (elect_one()) mbar_arrive(fb + 8 * s + eoff);   // both rows of the stage are in registers

                const bool emit = t >= (K - 1) / 2;
                const bool two = ro + 1 < yb;
                uint32_t h[2][12];
#pragma unroll
                for (int rho = 0; rho < 2; ++rho) {
                    // planar words: X[ch][m] = channel ch of pixels 4m .. 4m+3 (pixels past NPX meet zero taps)
                    uint32_t X[3][NG];
#pragma unroll
                    for (int m = 0; m < NG; ++m) {
                        const uint32_t wa = w[rho][3 * m];
                        const uint32_t wb = 3 * m + 1 < NWRD ? w[rho][3 * m + 1] : 0u;
                        X[0][m] = prmt(wa, wb, 0x0630);     // r0 r1 r2 .
                        X[1][m] = prmt(wa, wb, 0x0741);     // g0 g1 g2 .
                        X[2][m] = prmt(wa, wb, 0x0052);     // b0 b1 . .
                        if (3 * m + 2 < NWRD) {
                            const uint32_t wc = w[rho][3 * m + 2];
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
                if (!VBOX) {
#pragma unroll
                    for (int c = 0; c < 12; ++c) W[s * 12 + c] = prmt(h[0][c], h[1][c], 0x5410);
                }

                // value j of the lane (byte column 12*lane + j) is pixel j/3, channel j%3: column index (j%3)*4 + j/3
#define CS_COL(j) (((j) % 3) * 4 + (j) / 3)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    // Y[c]: the quotient of column c, in byte YB of the register (u8 output) or shifted down (int32 output)
                    uint32_t Y[12];
                    constexpr int YB = VBOX ? (DIV == 1 ? 2 : 0) : DIV == 3 ? 1 : DIV == 4 ? 2 : 0;
                    if (VBOX) {
                        // window sums of the K newest rows: S += P[row] - P[row-K], two columns per IADD3
                        constexpr int R = K + 1;
                        const int slot = 2 * s + half, old = (2 * s + half + 1) % R;
#pragma unroll
                        for (int pc = 0; pc < 6; ++pc) {
                            const uint32_t pn = prmt(h[half][2 * pc], h[half][2 * pc + 1], 0x5410);
                            SUM[pc] = SUM[pc] + pn - W[old * 6 + pc];
                            W[slot * 6 + pc] = pn;
                        }
                        if (emit) {
#pragma unroll
                            for (int pc = 0; pc < 6; ++pc) {
                                if (DIV == 1) {
                                    // floor(s/norm) sits in bits 16.. of umulhi(s << 16 | anything below, magic) (host-checked range)
                                    Y[2 * pc + 1] = __umulhi(SUM[pc], p.magic);
                                    Y[2 * pc] = __umulhi(SUM[pc] << 16, p.magic);
                                    if (!OUT_U8) { Y[2 * pc + 1] >>= 16; Y[2 * pc] >>= 16; }
                                } else {
                                    Y[2 * pc + 1] = SUM[pc] >> (16 + p.shift);
                                    Y[2 * pc] = (SUM[pc] & 0xffffu) >> p.shift;
                                }
                            }
                        }
                    } else if (emit) {
#pragma unroll
                        for (int c = 0; c < 12; ++c) {
                            uint32_t acc = 0u;
#pragma unroll
                            for (int k = 0; k < NP; ++k) {
                                const uint32_t wq = W[((s + 1 + k) % NP) * 12 + c];
                                acc = half ? dp2<SG, true>(wq, p.ut[k], acc) : dp2<SG, false>(wq, p.ut[k], acc);
                            }
                            if (DIV == 0) Y[c] = acc >> p.shift;
                            else if (DIV == 1) Y[c] = __umulhi(acc, p.magic);
                            else if (DIV == 2) Y[c] = (uint32_t)((int32_t)acc / p.norm);
                            else Y[c] = OUT_U8 ? acc : acc >> (8 * YB);
                        }
                    }
                    if (emit) {
                        uint8_t *grow = orow + (size_t)half * p.out_pitch;
                        const bool live = half == 0 || two;
                        if (OUT_U8) {
                            constexpr uint32_t SEL = 0x0040u + 0x11u * YB;
                            uint32_t o[3];
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                o[k] = prmt(prmt(Y[CS_COL(4 * k)], Y[CS_COL(4 * k + 1)], SEL), prmt(Y[CS_COL(4 * k + 2)], Y[CS_COL(4 * k + 3)], SEL), 0x5410);
                            if (wide) {
                                const uint32_t d = stg + half * WROW;
                                sts32(d, o[0]); sts32(d + 4, o[1]); sts32(d + 8, o[2]);
                            } else if (lval > 0 && live) {
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
                            if (wide) {
                                const uint32_t d = stg + half * WROW;
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    sts128(d + 16 * k, make_uint4(Y[CS_COL(4 * k)], Y[CS_COL(4 * k + 1)], Y[CS_COL(4 * k + 2)], Y[CS_COL(4 * k + 3)]));
                            } else if (lval > 0 && live) {
                                uint32_t *gp = reinterpret_cast<uint32_t *>(grow) + lane * 12;
#pragma unroll
                                for (int j = 0; j < 12; ++j)
                                    if (j < lval) gp[j] = Y[CS_COL(j)];
                            }
                        }
                    }
                }
#undef CS_COL
                if (emit) {
                    if (wide) {
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < NPC; ++j) {
                            const int c = lane + 32 * j;
                            if (2 * PPR % 32 == 0 || j + 1 < NPC || c < 2 * PPR) {
                                const int row = c >= PPR;
                                if (row == 0 || two)
                                    *reinterpret_cast<uint4 *>(orow + (row ? p.out_pitch : 0) + (size_t)(c - row * PPR) * 16) = lds128(stg_rd + 512u * j);
                            }
                        }
                    }
                    orow += 2 * p.out_pitch;
                    ro += 2;
                }
            }
        }
        so += NP * stage_bytes;
        fb += 8 * NP;
        if (so == ring_bytes) { so = 0; fb = full0; ph ^= 1; }
    }
}

template <int K>
cudaError_t launch_conv_sep_k(const ConvSepParams &p, int out_u8, int div, int vbox, dim3 grid, int smem, cudaStream_t st) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        count_launch();
        kern<<<grid, (p.nw + 1) * 32, smem, st>>>(p);
        return cudaGetLastError();
    };
    if (vbox) {
        if (div == 0) return out_u8 ? go(conv_sep_kernel<K, true, 0, true>) : go(conv_sep_kernel<K, false, 0, true>);
        if (div == 1) return out_u8 ? go(conv_sep_kernel<K, true, 1, true>) : go(conv_sep_kernel<K, false, 1, true>);
        return cudaErrorInvalidValue;
    }
    switch (div) {
        case 0: return out_u8 ? go(conv_sep_kernel<K, true, 0, false>) : go(conv_sep_kernel<K, false, 0, false>);
        case 1: return out_u8 ? go(conv_sep_kernel<K, true, 1, false>) : go(conv_sep_kernel<K, false, 1, false>);
        case 2: return out_u8 ? go(conv_sep_kernel<K, true, 2, false>) : go(conv_sep_kernel<K, false, 2, false>);
        case 3: if (out_u8) return go(conv_sep_kernel<K, true, 3, false>); break;
        case 4: if (out_u8) return go(conv_sep_kernel<K, true, 4, false>); break;
    }
    return cudaErrorInvalidValue;
}

#endif  // OMPIMG_CONV_SEP_IMPL

}  // namespace ompimg
