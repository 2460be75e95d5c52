// conv_sep.cuh -- K x K convolution with rank-1 taps (box, binomial/Gaussian, ...) on interleaved
// uint8 RGB: the fast path of a6 (old/parallel_convolution.c:8-24).  Same results as conv.cu's
// kernels: out[r][c][p] = (sum_{kr,kc} in[r+kr][c+kc][p] * u[kr]*v[kc]) / norm, C division, no clamp,
// pre-padded top-left anchored input.  Instantiated by conv_sep_inst.cu, once per K and group of variants.
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
// Arithmetic per lane and input row, and the ranges the host checks: conv_lane.cuh (shared with conv_tma.cuh).
#pragma once
#include "common.cuh"
#include "conv_lane.cuh"
#include "kernels.h"

namespace ompimg {

#ifndef CS_MINB5
#define CS_MINB5 4                     // resident CTAs per SM the K <= 5 kernels are compiled for (K > 5: 3)
#endif
#ifndef CS_MINB9
#define CS_MINB9 3
#endif
#ifndef CS_NWARPS
#define CS_NWARPS 4
#endif
constexpr int CS_NW = CS_NWARPS;         // consumer warps per CTA
constexpr int CS_LPX = CL_LPX;           // output pixels per lane
constexpr int CS_WPX = CL_WPX;           // output pixels per warp and row

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

// The kernel variants of one K are instantiated in eight translation units (conv_sep_inst.cu compiled once per K and PART, so that they
// build in parallel): PART 0 / 1 constant vertical taps with int32 / uint8 output, 2 / 3 general taps with int32 / uint8 output, 4..7 the
// same four for unit-ended taps (ConvLane's E1).
template <int K, int PART>
cudaError_t launch_conv_sep_part(const ConvSepParams &p, int div, dim3 grid, int smem, cudaStream_t st);
template <int K>
inline cudaError_t launch_conv_sep_k(const ConvSepParams &p, int out_u8, int div, int vbox, int ends1, dim3 grid, int smem, cudaStream_t st) {
    if (ends1 && div != 2) {
        if (vbox) return out_u8 ? launch_conv_sep_part<K, 5>(p, div, grid, smem, st) : launch_conv_sep_part<K, 4>(p, div, grid, smem, st);
        return out_u8 ? launch_conv_sep_part<K, 7>(p, div, grid, smem, st) : launch_conv_sep_part<K, 6>(p, div, grid, smem, st);
    }
    if (vbox) return out_u8 ? launch_conv_sep_part<K, 1>(p, div, grid, smem, st) : launch_conv_sep_part<K, 0>(p, div, grid, smem, st);
    return out_u8 ? launch_conv_sep_part<K, 3>(p, div, grid, smem, st) : launch_conv_sep_part<K, 2>(p, div, grid, smem, st);
}

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
// DIV: 0 acc >> shift | 1 umulhi(acc, magic) | 2 acc / norm (signed) | 3 byte 1 (norm 256, u8 out) | 4 byte 2 (norm 65536, u8 out)
// VBOX (DIV 0 or 1 only): see the header comment.
template <int K, bool OUT_U8, int DIV, bool VBOX, bool E1 = false>
__global__ void __launch_bounds__((CS_NW + 1) * 32, K <= 5 ? CS_MINB5 : CS_MINB9) conv_sep_kernel(const __grid_constant__ ConvSepParams p) {
    using Lane = ConvLane<K, OUT_U8, DIV, VBOX, E1>;
    constexpr int NP = Lane::NP;                    // trips per unrolled body; pair-rows in the vertical window
    constexpr int NWRD = Lane::NWRD;                // 32-bit words that hold a lane's 4+K-1 input pixels
    constexpr int OB = Lane::OB;
    constexpr uint32_t WROW = Lane::WROW;           // staging bytes per warp and output row
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
    // staging: two output rows per warp (conv_lane.cuh)
    const uint32_t stg_warp = staging0 + (uint32_t)warp * (2u * WROW);
    const uint32_t stg = stg_warp + (uint32_t)lane * (12u * OB);
    const uint32_t stg_rd = stg_warp + (uint32_t)lane * 16u;
    uint32_t mrow = (uint32_t)(uintptr_t)p.in + (uint32_t)ya * pitch_lo;        // low address bits of the current input row
    uint8_t *orow = p.out + (size_t)ya * p.out_pitch + (size_t)seg0 * OB;       // next output row of this warp
    int ro = ya;

    Lane L;                      // the vertical window (register ring / running sums) of this lane's 12 output columns
    L.init();

    uint32_t so = 0, fb = full0, ph = 0;           // ring position of the unrolled body's first trip
    int t0 = 0;
    // one trip (two input rows) at ring slot s of the unrolled body.  CHECKED: the trip may not exist, may not emit (priming), may hold
    // one output row only, and the warp's output may need per-lane stores; the unchecked body has none of that code (conv_tma.cuh).
    auto trip = [&](auto s_c, auto checked_c) __attribute__((always_inline)) {
        constexpr int s = decltype(s_c)::value;
        constexpr bool CHECKED = decltype(checked_c)::value;
        const int t = t0 + s;
        if (CHECKED && t >= ntrips) return;
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

        const bool emit = !CHECKED || t >= NP - 1;
        const bool two = !CHECKED || ro + 1 < yb;
        L.template rows2<s, !CHECKED>(w, p, emit, two, wide, lval, lane, stg, stg_rd, orow, p.out_pitch);
        if (emit) {
            orow += 2 * p.out_pitch;
            ro += 2;
        }
    };
    auto block = [&](auto checked_c) __attribute__((always_inline)) {
        trip(ClInt<0>{}, checked_c);
        if constexpr (NP > 1) trip(ClInt<1>{}, checked_c);
        if constexpr (NP > 2) trip(ClInt<2>{}, checked_c);
        if constexpr (NP > 3) trip(ClInt<3>{}, checked_c);
        if constexpr (NP > 4) trip(ClInt<4>{}, checked_c);
        t0 += NP;
        so += NP * stage_bytes;
        fb += 8 * NP;
        if (so == ring_bytes) { so = 0; fb = full0; ph ^= 1; }
    };
    // block 0 primes the window; blocks whose trips all exist and all emit two rows run unchecked when the warp's output leaves as
    // whole 16-byte pieces; the rest checked
    const int tfull = ((yb - ya) >> 1) + NP - 1;   // trips [NP-1, tfull) emit two rows each
    block(ClBool<true>{});
    while (wide && t0 + NP <= tfull) block(ClBool<false>{});
    while (t0 < ntrips) block(ClBool<true>{});
}

template <int K, int PART>
cudaError_t launch_conv_sep_part(const ConvSepParams &p, int div, dim3 grid, int smem, cudaStream_t st) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        count_launch();
        kern<<<grid, (p.nw + 1) * 32, smem, st>>>(p);
        return cudaGetLastError();
    };
    constexpr bool U8 = (PART & 1) != 0, VB = (PART & 2) == 0, E1 = PART >= 4;        // PART 4..7: unit-ended taps (ConvLane's E1), unsigned only
    if constexpr (VB) {                   // constant vertical taps
        if (div == 0) return go(conv_sep_kernel<K, U8, 0, true, E1>);
        if (div == 1) return go(conv_sep_kernel<K, U8, 1, true, E1>);
        return cudaErrorInvalidValue;
    } else {
        switch (div) {
            case 0: return go(conv_sep_kernel<K, U8, 0, false, E1>);
            case 1: return go(conv_sep_kernel<K, U8, 1, false, E1>);
            case 2: if constexpr (!E1) return go(conv_sep_kernel<K, U8, 2, false>); break;
            case 3: if constexpr (U8) return go(conv_sep_kernel<K, true, 3, false, E1>); break;
            case 4: if constexpr (U8) return go(conv_sep_kernel<K, true, 4, false, E1>); break;
        }
        return cudaErrorInvalidValue;
    }
}

#endif  // OMPIMG_CONV_SEP_IMPL

}  // namespace ompimg
