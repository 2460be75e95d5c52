// conv_lane.cuh -- the per-lane arithmetic of the rank-1 K x K convolution (a6: old/parallel_convolution.c:8-24), shared by the
// two kernels that differ only in how rows reach shared memory: conv_tma.cuh (warp-autonomous, tensor-map copies) and
// conv_sep.cuh (producer warp, 1-D bulk copies).  A lane owns 4 pixels (12 output bytes) of a 128-pixel warp strip and is handed
// two consecutive input rows per call (its 4+K-1 pixels of each, as 32-bit words):
//   1. de-interleave: 2 PRMT per planar word (r0 r1 r2 r3), (g0..g3), (b0..b3) -- this removes the dilation-by-3 that
//      interleaved channels impose on the horizontal pass;
//   2. horizontal pass: h[i] = sum_m dp4a(X[m], TH[i][m]) with the taps v pre-shifted into byte position on the host
//      (TH[i][m].byte b = v[4m+b-i]): ceil((K+3)/4) IDP.4A per value, no shifts;
//   3. vertical pass, two forms:
//      general u (VBOX = false): the two rows' h are packed as u16 pairs (one PRMT) and
//        out[ro] = sum_q dp2a(W[q], UT[q]) over the (K+1)/2 newest pair-rows, kept in a register ring whose slot S is a
//        compile-time argument (the callers unroll (K+1)/2 calls);
//      constant u (VBOX = true, box blur): two neighbouring columns' h are packed as u16 pairs and a running window sum
//        S += P[r] - P[r-K] (ONE IADD3 per pair and row, both halves at once; the last K+1 rows of P stay in the ring)
//        replaces the K multiply-adds;
//   4. division by the norm: a shift, one multiply-high (exact: the host checks the range), a byte select folded into the
//      pack (norm 256 / 65536), or a true division (signed taps);
//   5. output: the two rows are transposed through a per-warp staging buffer and leave as 128-bit stores of contiguous rows
//      (`wide`), or as per-lane stores where the strip is ragged or the output rows are not 16-byte aligned.
// Ranges (checked on the host, conv_sep.cu): unsigned form: taps >= 0, v,u <= 255, 255*sum(v) <= 65535 (h fits u16); signed form
// (DIV 2): |v|,|u| <= 127, 255*sum|v| <= 32767; both: 255*sum|u|*sum|v| < 2^31; VBOX: the window sum fits u16 as well.
#pragma once
#include "common.cuh"

namespace ompimg {

constexpr int CL_LPX = 4;                // output pixels per lane
constexpr int CL_WPX = 32 * CL_LPX;      // output pixels per warp and row
template <int N> struct ClInt { static constexpr int value = N; };       // compile-time ring slot for the callers' unrolled bodies
template <bool B> struct ClBool { static constexpr bool value = B; };    // checked / unchecked body

// DIV: 0 acc >> shift | 1 umulhi(acc, magic) | 2 acc / norm (signed taps) | 3 byte 1 (norm 256, u8 out) | 4 byte 2 (norm 65536, u8 out)
// E1 (unsigned taps): v[0] == v[K-1] == 1 and, unless VBOX, u[0] == u[K-1] == 1 (binomial and other unit-ended taps, the plain box).  A tap word of the horizontal pass
// that holds a single tap, and the half pair-row at either end of the vertical window, then contribute the sample itself: one byte /
// half-word extraction (ALU pipe) starts the accumulator instead of an IDP -- 1.5 of the 8 IDP per output byte at K = 9, on the pipe
// that bounds that kernel.
template <int K, bool OUT_U8, int DIV, bool VBOX, bool E1 = false>
struct ConvLane {
    static constexpr bool SG = DIV == 2;
    static constexpr int NP = (K + 1) / 2;                 // calls per unrolled block; pair-rows in the vertical window
    static constexpr int NPX = CL_LPX + K - 1;             // input pixels per lane
    static constexpr int NWRD = (NPX * 3 + 3) / 4;         // 32-bit words that hold them
    static constexpr int NG = (NPX + 3) / 4;               // planar words per channel
    static constexpr int OB = OUT_U8 ? 1 : 4;
    static constexpr uint32_t WROW = CL_WPX * 3 * OB;      // staging bytes per warp and output row
    static constexpr int RING = VBOX ? (K + 1) * 6 : NP * 12;
    // staging: two output rows per warp; written as 12 values per lane, read back as 16-byte pieces
    // (piece c = lane + 32*j lies in row c / PPR at byte (c % PPR) * 16; j < NPC)
    static constexpr int PPR = (int)WROW / 16, NPC = 2 * PPR / 32 + (2 * PPR % 32 ? 1 : 0);

    uint32_t W[RING];            // VBOX: P[slot][6] (slot = input row mod K+1); else W[pair-row slot][12]
    uint32_t SUM[6];             // VBOX: running window sums, two columns per register

    template <bool SGN>
    static __device__ __forceinline__ uint32_t dp4(uint32_t a, uint32_t b, uint32_t c) {
        uint32_t d;
        if (SGN) asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        else asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        return d;
    }
    template <bool SGN, bool HI>
    static __device__ __forceinline__ uint32_t dp2(uint32_t a, uint32_t b, uint32_t c) {
        uint32_t d;
        if (SGN) {
            if (HI) asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
            else asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        } else {
            if (HI) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
            else asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        }
        return d;
    }

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int q = 0; q < RING; ++q) W[q] = 0u;
#pragma unroll
        for (int q = 0; q < 6; ++q) SUM[q] = 0u;
    }

    // Two consecutive input rows (w[0], w[1]) at ring slot S.  p: the kernel's parameter block (th, ut, magic, shift, norm).
    // emit: the rows complete output rows (orow, orow + opitch); two: the second of them exists.  stg / stg_rd: this lane's
    // write / read address in the warp's staging buffer.  The caller advances orow by two rows when emit.
    // WIDE_ONLY: the caller guarantees `wide` (its steady loop): the per-lane store code is not generated at all, which keeps the
    // unrolled hot block small (instruction fetch is a limit for the K = 9 bodies).
    template <int S, bool WIDE_ONLY = false, typename P>
    __device__ __forceinline__ void rows2(const uint32_t (&w)[2][NWRD], const P &p, bool emit, bool two, bool wide_in, int lval, int lane,
                                          uint32_t stg, uint32_t stg_rd, uint8_t *orow, size_t opitch) {
        const bool wide = WIDE_ONLY ? true : wide_in;
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
                    if (E1) {
#pragma unroll
                        for (int m = 0; m < NG; ++m) {          // words that hold exactly one tap (an end tap, = 1): add the sample byte
                            const int first = 4 * m - i, last = 4 * m + 3 - i;
                            const int lo = first < 0 ? 0 : first, hi = last > K - 1 ? K - 1 : last;
                            if (hi == lo && hi >= 0 && lo <= K - 1) acc += (X[ch][m] >> (8 * (lo - first))) & 0xffu;
                        }
                    }
#pragma unroll
                    for (int m = 0; m < NG; ++m) {
                        const int first = 4 * m - i, last = 4 * m + 3 - i;
                        const int lo = first < 0 ? 0 : first, hi = last > K - 1 ? K - 1 : last;
                        if (hi >= lo && !(E1 && hi == lo)) acc = dp4<SG>(X[ch][m], p.th[i][m < 3 ? m : 0], acc);
                    }
                    h[rho][ch * 4 + i] = acc;
                }
        }
        if (!VBOX) {
#pragma unroll
            for (int c = 0; c < 12; ++c) W[S * 12 + c] = prmt(h[0][c], h[1][c], 0x5410);
        }

        // value j of the lane (byte column 12*lane + j) is pixel j/3, channel j%3: column index (j%3)*4 + j/3
#define CL_COL(j) (((j) % 3) * 4 + (j) / 3)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            // Y[c]: the quotient of column c, in byte YB of the register (u8 output) or shifted down (int32 output)
            uint32_t Y[12];
            constexpr int YB = VBOX ? (DIV == 1 ? 2 : 0) : DIV == 3 ? 1 : DIV == 4 ? 2 : 0;
            if (VBOX) {
                // window sums of the K newest rows: S += P[row] - P[row-K], two columns per IADD3
                constexpr int R = K + 1;
                constexpr int slot_base = 2 * S;
                const int slot = slot_base + half, old = (slot_base + half + 1) % R;
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
                    // E1: the newest pair-row holds only the window's last row for the first output row (tap u[K-1] = 1, low half),
                    // the oldest only its first row for the second output row (tap u[0] = 1, high half)
                    if (E1) acc = half ? W[((S + 1) % NP) * 12 + c] >> 16 : W[S * 12 + c] & 0xffffu;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        if (E1 && k == (half ? 0 : NP - 1)) continue;
                        const uint32_t wq = W[((S + 1 + k) % NP) * 12 + c];
                        acc = half ? dp2<SG, true>(wq, p.ut[k], acc) : dp2<SG, false>(wq, p.ut[k], acc);
                    }
                    if (DIV == 0) Y[c] = acc >> p.shift;
                    else if (DIV == 1) Y[c] = __umulhi(acc, p.magic);
                    else if (DIV == 2) Y[c] = (uint32_t)((int32_t)acc / p.norm);
                    else Y[c] = OUT_U8 ? acc : acc >> (8 * YB);
                }
            }
            if (emit) {
                uint8_t *grow = orow + (size_t)half * opitch;
                const bool live = half == 0 || two;
                if (OUT_U8) {
                    constexpr uint32_t SEL = 0x0040u + 0x11u * YB;
                    uint32_t o[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        o[k] = prmt(prmt(Y[CL_COL(4 * k)], Y[CL_COL(4 * k + 1)], SEL), prmt(Y[CL_COL(4 * k + 2)], Y[CL_COL(4 * k + 3)], SEL), 0x5410);
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
                            sts128(d + 16 * k, make_uint4(Y[CL_COL(4 * k)], Y[CL_COL(4 * k + 1)], Y[CL_COL(4 * k + 2)], Y[CL_COL(4 * k + 3)]));
                    } else if (lval > 0 && live) {
                        uint32_t *gp = reinterpret_cast<uint32_t *>(grow) + lane * 12;
#pragma unroll
                        for (int j = 0; j < 12; ++j)
                            if (j < lval) gp[j] = Y[CL_COL(j)];
                    }
                }
            }
        }
#undef CL_COL
        if (emit) {
            if (wide) {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NPC; ++j) {
                    const int c = lane + 32 * j;
                    if (2 * PPR % 32 == 0 || j + 1 < NPC || c < 2 * PPR) {
                        const int row = c >= PPR;
                        if (row == 0 || two)
                            *reinterpret_cast<uint4 *>(orow + (row ? opitch : 0) + (size_t)(c - row * PPR) * 16) = lds128(stg_rd + 512u * j);
                    }
                }
            }
        }
    }
};

}  // namespace ompimg
