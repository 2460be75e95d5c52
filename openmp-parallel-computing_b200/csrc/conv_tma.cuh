// conv_tma.cuh -- K x K convolution with rank-1 taps on interleaved uint8 RGB, warp-autonomous form:
// the fast path of a6 (old/parallel_convolution.c:8-24) for inputs whose base and pitch are 16-byte
// aligned (every device copy the host entry points make; conv_sep.cuh keeps the dense-pitch layouts).
// Same results as conv.cu / conv_sep.cuh: out[r][c][p] = (sum_{kr,kc} in[r+kr][c+kc][p] * u[kr]*v[kc]) / norm,
// C division, no clamp, pre-padded top-left anchored input.  Instantiated by conv_tma_inst.cu, once per K and group of variants.
//
// Data movement.  Every WARP is its own pipeline (as in stencil.cu): it owns a band of output rows x a
// strip of 128 pixels, runs a private ring of two-row stages with one mbarrier per stage, and an elected
// lane requests a whole stage with ONE cp.async.bulk.tensor.2d (SASS UTMALDG): box = {(128+K-1)*3 bytes
// rounded up to 16, 2 rows} of the tensor {row bytes, rows+K-1}.  Parts of a box right of the row or
// below the image are zero-filled by the copy engine, so the kernel holds no address arithmetic for its
// loads, no clamping and no dummy rows.  There is no producer warp, no empty barrier and no CTA-wide
// barrier (conv_sep.cuh: 4.0 % SYNCS + 3.4 % NANOSLEEP of all instructions, a fifth of the CTA's warps,
// and consumers that wait for the slowest of four warps to release a stage).
//
// Loop structure.  A band's trips (two input rows each) run in blocks of NP = (K+1)/2 trips, so that the
// ring slot of the vertical window is a compile-time register.  The first block (K-1 priming rows, first
// outputs) and the last trips of the band run a "checked" body; every block in between runs a body with
// no per-trip tests (emit, second-row-valid, trip-exists) and, being entered only by warps whose output leaves as whole
// 16-byte pieces, no per-lane store code.
//
// Arithmetic per lane (4 pixels = 12 output bytes) and input row: conv_lane.cuh (shared with conv_sep.cuh).
// Instruction budget and the pipe that bounds each K: DESIGN.md 4.3.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_lane.cuh"
#include "kernels.h"

namespace ompimg {

// resident CTAs per SM the kernels are compiled for (launch bounds -> register budget).  Measured (tools/conv_ab.sh): the 9x9 kernels
// want registers rather than warps -- box 676 Gpx/s at 4 CTAs, 722 at 3; binomial 502 at 5 (tiny spills), 553 at 4, 591 at 3 --
// the 5x5 box 811 at 5, 831 at 4, the 5x5 binomial 824 at 5, 817 at 4.
#ifndef CT_MINB_5BOX
#define CT_MINB_5BOX 4
#endif
#ifndef CT_MINB_5GEN
#define CT_MINB_5GEN 5
#endif
#ifndef CT_MINB_9BOX
#define CT_MINB_9BOX 3
#endif
#ifndef CT_MINB_9GEN
#define CT_MINB_9GEN 3
#endif
#ifndef CT_E1_MIN_K
#define CT_E1_MIN_K 3                     // smallest K that gets the unit-ended-taps variant (5x5 binomial 824 -> 898 Gpx/s, 9x9 591 -> 674)
#endif
constexpr int CT_WARPS = 4;              // warps (tasks) per CTA
constexpr int CT_LPX = CL_LPX;           // output pixels per lane
constexpr int CT_WPX = CL_WPX;           // output pixels per warp and row

constexpr uint32_t ct_box_bytes(int K) { return (uint32_t)(((CT_WPX + K - 1) * 3 + 15) & ~15); }
constexpr uint32_t ct_stage_bytes(int K) { return (2 * ct_box_bytes(K) + 127u) & ~127u; }

struct ConvTmaParams {
    CUtensorMap map;             // {ceil(row bytes / 8), rows+K-1} of uint64 elements, box {box bytes / 8, 2}
    uint8_t *out;
    size_t out_pitch;
    int rows, cols;
    int band_rows;               // even
    int nstrips;
    long long ntasks;            // nbands * nstrips, one task per warp
    int stages;                  // ring depth per warp (>= 2)
    uint32_t out_wide;           // output rows are 16-byte aligned: 128-bit stores through the staging buffer
    uint32_t th[4][3];           // horizontal tap words
    uint32_t ut[5];              // vertical tap-pair words: bytes (u[2q], u[2q+1], u[2q-1], u[2q])
    uint32_t magic;
    int32_t shift;
    int32_t norm;                // unused here (signed taps take conv_sep.cuh); read by the shared lane code's DIV 2 branch only
};

// band geometry and ring depth are chosen by the launcher from the kernel's occupancy (conv_sep.cu passes everything else).
// The kernel variants of one K are instantiated in four translation units (conv_tma_inst.cu compiled once per K and PART, so that
// they build in parallel): PART 0 constant vertical taps (box), 1 the same with unit-ended horizontal taps, 2 general taps, 3 general
// unit-ended taps.
template <int K, int PART>
cudaError_t launch_conv_tma_part(ConvTmaParams &p, int out_u8, int div, int sm_count, cudaStream_t st);
template <int K>
inline cudaError_t launch_conv_tma_k(ConvTmaParams &p, int out_u8, int div, int vbox, int ends1, int sm_count, cudaStream_t st) {
    const bool e1 = ends1 && K >= CT_E1_MIN_K;
    if (vbox) return e1 ? launch_conv_tma_part<K, 1>(p, out_u8, div, sm_count, st) : launch_conv_tma_part<K, 0>(p, out_u8, div, sm_count, st);
    return e1 ? launch_conv_tma_part<K, 3>(p, out_u8, div, sm_count, st) : launch_conv_tma_part<K, 2>(p, out_u8, div, sm_count, st);
}

#ifdef OMPIMG_CONV_TMA_IMPL

__device__ __forceinline__ bool ct_elect() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// DIV: 0 acc >> shift | 1 umulhi(acc, magic) | 3 byte 1 (norm 256, u8 out) | 4 byte 2 (norm 65536, u8 out); taps >= 0.
// VBOX (DIV 0 or 1 only): constant vertical taps, running window sums.
template <int K, bool OUT_U8, int DIV, bool VBOX, bool E1 = false>
__global__ void __launch_bounds__(CT_WARPS * 32, K <= 5 ? (VBOX ? CT_MINB_5BOX : CT_MINB_5GEN) : (VBOX ? CT_MINB_9BOX : CT_MINB_9GEN)) conv_tma_kernel(const __grid_constant__ ConvTmaParams p) {
    using Lane = ConvLane<K, OUT_U8, DIV, VBOX, E1>;
    constexpr int NP = Lane::NP;                    // trips per block; pair-rows in the vertical window
    constexpr int NWRD = Lane::NWRD;                // 32-bit words that hold a lane's 4+K-1 input pixels
    constexpr int OB = Lane::OB;
    constexpr uint32_t WROW = Lane::WROW;           // staging bytes per output row
    constexpr uint32_t BOXB = ct_box_bytes(K), STAGE = ct_stage_bytes(K);
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);           // warp-uniform by construction
    const long long task = (long long)blockIdx.x * CT_WARPS + warp;
    if (task >= p.ntasks) return;                   // warps are independent: no CTA-wide barrier below
    // strip-major inside a band: the warps of a CTA read one contiguous piece of every row
    const int band = (int)(task / p.nstrips), strip = (int)(task - (long long)band * p.nstrips);
    const int S = p.stages;
    const uint32_t warp_bytes = (uint32_t)S * STAGE + 2u * WROW;
    const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
    const uint32_t ring0 = base + (uint32_t)warp * warp_bytes;
    const uint32_t staging0 = ring0 + (uint32_t)S * STAGE;
    const uint32_t bar0 = base + CT_WARPS * warp_bytes + (uint32_t)warp * (8u * (uint32_t)S);

    const int xs = strip * CT_WPX;
    const int ya = band * p.band_rows, yb = min(p.rows, ya + p.band_rows);
    const int nout = yb - ya;
    const int ntrips = (nout + K) >> 1;             // input rows ya .. yb+K-2, two per trip (a missing last row is zero-filled)

    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int cx = xs * 3 / 8;                      // box start in uint64 elements (384 bytes per strip)
    auto request = [&](uint32_t stage_off, uint32_t bar, int tt) {                 // elected lane only
        mbar_expect_tx(bar, 2u * BOXB);
        tma_load_2d(ring0 + stage_off, &p.map, cx, ya + 2 * tt, bar);
    };
    if (ct_elect()) {
        for (int t = 0; t < S && t < ntrips; ++t) request((uint32_t)t * STAGE, bar0 + 8 * t, t);
    }

    const int seg0 = xs * 3;                                                    // first output byte column of the warp
    const int nval = max(0, min(CT_WPX * 3, p.cols * 3 - seg0));                 // valid output byte columns of the warp
    const bool wide = p.out_wide && nval == CT_WPX * 3;
    const int lval = nval - lane * 12;                                          // valid output values of the lane (<= 0: none)
    const uint32_t my = keep(ring0 + (uint32_t)lane * (CT_LPX * 3u));
    const uint32_t ring_bytes = (uint32_t)S * STAGE;
    const uint32_t stg = staging0 + (uint32_t)lane * (12u * OB);
    const uint32_t stg_rd = staging0 + (uint32_t)lane * 16u;
    const size_t opitch = p.out_pitch;
    uint8_t *orow = p.out + (size_t)ya * opitch + (size_t)seg0 * OB;            // next output row of this warp
    int ro = ya;

    Lane L;                      // the vertical window (register ring / running sums) of this lane's 12 output columns
    L.init();

    uint32_t so = 0, fb = bar0, ph = 0;            // ring position (stage offset, its barrier, its phase) of the next trip
    int t0 = 0;                                    // trip number of the current block's first trip

    // one trip: slot s_c of the current block.  CHECKED: the trip may not exist, may not emit (priming), may hold one output row only.
    auto trip = [&](auto s_c, auto checked_c) __attribute__((always_inline)) {
        constexpr int s = decltype(s_c)::value;
        constexpr bool CHECKED = decltype(checked_c)::value;
        const int t = t0 + s;
        if (CHECKED && t >= ntrips) return;
        mbar_wait(fb, ph);
        uint32_t w[2][NWRD];
#pragma unroll
        for (int rho = 0; rho < 2; ++rho) {
            const uint32_t a = my + so + (uint32_t)rho * BOXB;
#pragma unroll
            for (int i = 0; i < NWRD; ++i) w[rho][i] = lds32(a + 4 * i);
        }
        // The refill below is an async-proxy WRITE of the stage these generic-proxy loads READ: the proxy fence orders each
        // lane's loads before it, the warp barrier collects the lanes.  Without it the copy (an L2 hit on small images) can land
        // before loads that still wait in the LSU queue behind this lane's output stores (seen with int32 per-lane stores).
        fence_async_smem();
        __syncwarp();                                  // every lane holds both rows (and has read the previous trip's staging rows)
        if (ct_elect() && t + S < ntrips) request(so, fb, t + S);
        so += STAGE; fb += 8;
        if (so == ring_bytes) { so = 0; fb = bar0; ph ^= 1; }

        const bool emit = !CHECKED || t >= NP - 1;
        const bool two = !CHECKED || ro + 1 < yb;
        L.template rows2<s, !CHECKED>(w, p, emit, two, wide, lval, lane, stg, stg_rd, orow, opitch);      // unchecked blocks run for `wide` warps only
        if (emit) {
            orow += 2 * opitch;
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
    };

    // block 0 primes the window; blocks whose trips all exist and all emit two rows run unchecked; the rest checked
    const int tfull = (nout >> 1) + NP - 1;        // trips [NP-1, tfull) emit two rows each
    block(ClBool<true>{});
    while (wide && t0 + NP <= tfull) block(ClBool<false>{});      // ragged strips / unaligned output rows stay on the checked body
    while (t0 < ntrips) block(ClBool<true>{});
}

template <int K, int PART>
cudaError_t launch_conv_tma_part(ConvTmaParams &p, int out_u8, int div, int sm_count, cudaStream_t st) {
    auto go = [&](auto kern) -> cudaError_t {
        const int ob = out_u8 ? 1 : 4;
        static int stages_env = -1, band_env = -1, occ_env = -1;
        if (occ_env < 0) { const char *e = getenv("OMPIMG_CONV_OCC"); occ_env = e ? atoi(e) : 0; }       // tuning: cap the resident CTAs per SM
        if (stages_env < 0) { const char *e = getenv("OMPIMG_CONV_STAGES"); stages_env = e ? atoi(e) : 0; }
        if (band_env < 0) { const char *e = getenv("OMPIMG_CONV_BAND"); band_env = e ? atoi(e) : 0; }
        // ring depth: three two-row stages per warp measured best for every K (deeper rings are slower: tools/conv_sweep2.sh)
        p.stages = stages_env > 0 ? max(2, min(16, stages_env)) : 3;
        int smem = CT_WARPS * (p.stages * (int)ct_stage_bytes(K) + 2 * CT_WPX * 3 * ob + 8 * p.stages) + 128;
        if (occ_env > 0) smem = max(smem, (int)((227 * 1024) / occ_env - 1024) & ~127);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CT_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorInvalidConfiguration;
        // Band height (even).  A launch of T = bands x strips tasks runs in F = ceil(T / slots) waves on the resident warp slots.  With
        // one wave the launch lasts as long as its tallest band plus the K-1 priming rows every band re-reads and re-filters; with
        // many waves early finishers pick up new tasks and only the total work W (rows + K-1 summed over the tasks) counts.  Estimate
        // W/slots + (F*(c+K-1) - W/slots)/F, exact at F = 1 and -> W/slots for large F, and take the height that minimises it (ties:
        // the taller band).
        const double slots = (double)sm_count * occ * CT_WARPS;
        int rpb = 2;
        double best = -1.0;
        for (int c = 512; c >= 2 * K; c -= 2) {
            const long long nb = (p.rows + c - 1) / c;
            const double tasks = (double)nb * p.nstrips, full = (double)(long long)((tasks + slots - 1) / slots);
            const double work = (double)p.nstrips * ((double)p.rows + (double)(K - 1) * nb) / slots;
            const double span = work + (full * (c + K - 1) - work) / full;
            if (best < 0 || span < best * 0.999) { best = span; rpb = c; }
        }
        if (p.rows < 2 * K) rpb = (p.rows + 1) & ~1;
        if (band_env > 0) rpb = max(2, band_env + (band_env & 1));
        p.band_rows = rpb;
        const long long nbands = (p.rows + rpb - 1) / rpb;
        p.ntasks = nbands * p.nstrips;
        const long long ctas = (p.ntasks + CT_WARPS - 1) / CT_WARPS;
        if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
        count_launch();
        kern<<<(unsigned)ctas, CT_WARPS * 32, smem, st>>>(p);
        return cudaGetLastError();
    };
    constexpr bool VB = PART < 2, E1 = (PART & 1) != 0;
    if constexpr (E1 && K < CT_E1_MIN_K) return cudaErrorInvalidValue;
    else if constexpr (VB) {
        if (div == 0) return out_u8 ? go(conv_tma_kernel<K, true, 0, true, E1>) : go(conv_tma_kernel<K, false, 0, true, E1>);
        if (div == 1) return out_u8 ? go(conv_tma_kernel<K, true, 1, true, E1>) : go(conv_tma_kernel<K, false, 1, true, E1>);
        return cudaErrorInvalidValue;
    } else {
        switch (div) {
            case 0: return out_u8 ? go(conv_tma_kernel<K, true, 0, false, E1>) : go(conv_tma_kernel<K, false, 0, false, E1>);
            case 1: return out_u8 ? go(conv_tma_kernel<K, true, 1, false, E1>) : go(conv_tma_kernel<K, false, 1, false, E1>);
            case 3: if (out_u8) return go(conv_tma_kernel<K, true, 3, false, E1>); break;
            case 4: if (out_u8) return go(conv_tma_kernel<K, true, 4, false, E1>); break;
        }
        return cudaErrorInvalidValue;
    }
}

#endif  // OMPIMG_CONV_TMA_IMPL

}  // namespace ompimg
