// stencil.cu -- 3x3 Sobel (a2: monolithic/src/sobel.c:6-25) and the fused RGB -> luma -> Sobel path
// (a3..a5: monolithic/src/main_with_sobel.c:51-74), as one row-streaming kernel fed by tensor-map copies.
#include <cstdlib>

#include <cuda.h>      // CUtensorMap and its enums (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "kernels.h"
#include "tma.h"

namespace ompimg {

// ---------------------------------------------------------------------------------------------
// sobel_tma_kernel: every warp is its own pipeline, fed by 3-D tensor-map copies.
//
// A warp owns a 512-pixel column strip of a band of rows of one frame ("task").  It runs a private
// ring of stages -- three source rows each -- with one mbarrier per stage; an elected lane requests a
// whole stage with ONE cp.async.bulk.tensor (SASS UTMALDG): box = {strip bytes + a 16-byte halo block
// each side, 3 rows, 1 frame} of the tensor {row bytes, height, frames}.  Out-of-range parts of a box
// (left of column 0, right of the row, below the image) are zero-filled by the copy engine, so there
// is no address arithmetic, no clamping and no per-row request in the kernel.  No producer warp, no
// empty barriers, no CTA-wide barrier: the eight warps of a CTA never wait for each other.
//
// Arithmetic per row and lane (16 pixels): see common.cuh -- f16x2 row partials, two previous rows'
// partials kept in registers and rotated through three registers per pixel pair (three rows per loop
// trip, so no register moves), FHFMA squares, MUFU.SQRT, saturating pack.  Each source byte is read
// from HBM once and from shared memory once; the luma plane of the fused path never exists in memory.
//
// Two box layouts: `full` (one box per stage; plane and RGB) and `split` (two half-strip boxes per
// stage, each with its own halo blocks).  split serves (a) RGBA, whose full-strip box would exceed the
// 256-element box limit, and (b) "folded" tasks: a last strip of <= 256 pixels (3840 = 7.5 strips)
// would leave half a warp idle, so ONE warp takes that strip for TWO consecutive bands, 16 lanes each.
//
// Loop structure: the first trip (two priming rows) and the trips that may cross the band's end run a
// "checked" body; every trip in between runs a body without per-row tests.
//
// Row-band mode (HALO, ompimg_bands_* / _band_dev): source rows -1 and `height` live in another
// allocation, normally a peer GPU's memory (same process or CUDA-IPC mapped).  The elected lane
// fetches them with 1-D bulk copies (UBLKCP) straight over NVLink into two spare row slots, on the
// same mbarrier as the stage they belong to, and the consumer reads those two rows from the spare
// slots: the halo exchange is part of the kernel, there is no staging copy and no collective.
//
// In-place RGB output (MODE 2, ompimg_gray_sobel_pass_dev: one iteration of main_with_sobel.c:51-74 on the image itself).
// The kernel writes the replicated edge value over the RGB pixels it read (a4 fused as an epilogue: 3 x 128-bit stores
// per lane and row, 6 bytes/pixel moved instead of 8 with a separate replicate pass).  Being in place, a warp may
// overwrite pixels another warp still has to read: the 16-byte halo blocks either side of its strip and the first and
// last source row of its band.  pass_prepare_kernel therefore saves, before this kernel starts, (a) every band's first
// and last source row and (b) the two pixels either side of every strip boundary of every row (16 bytes per boundary, with
// the image-border values of the row in the outermost entries).  The kernel takes (a) into the spare row slots like
// row-band mode does and (b) with one more 96-byte tensor box per stage, and never looks at the halo blocks of its stages.
// ---------------------------------------------------------------------------------------------
#ifndef SOBEL_MINB_PLANE
#define SOBEL_MINB_PLANE 3                  // resident CTAs per SM the plane kernel is compiled for (80 registers): 2547 Gpx/s on 4K planes; 2 (96 registers) 2352, 4 (64, spills) 2058
#endif
#ifndef SOBEL_MINB_RGB
#define SOBEL_MINB_RGB 2                    // ... and the RGB(A) kernels (128 registers; shared memory allows no more): fused 1432 Gpx/s, 1 CTA 1378
#endif
constexpr int TW_WARPS = 8;
constexpr uint32_t round128(uint32_t v) { return (v + 127u) & ~127u; }
template <int CH>
struct TmaGeom {
    static constexpr int S = CH == 1 ? 4 : 2;                        // stages per warp
    static constexpr uint32_t RS = 32 + 512 * CH;                    // bytes of a full-strip row in a stage (halo blocks included)
    static constexpr uint32_t HS = 32 + 256 * CH;                    // ... of a half-strip row
    static constexpr bool HAS_FULL = RS / 8 <= 256;                  // box dimensions are limited to 256 elements (of 8 bytes here)
    static constexpr uint32_t HALFSTAGE = round128(3 * HS);
    static constexpr uint32_t STAGE = HAS_FULL && round128(3 * RS) > 2 * HALFSTAGE ? round128(3 * RS) : 2 * HALFSTAGE;
    static constexpr uint32_t HALO_SLOT = round128(HAS_FULL && RS > 2 * HS ? RS : 2 * HS);
    // mode 0: plain; 1: row-band mode (two spare row slots); 2: in-place RGB output (spare row slots + a 128-byte side box per stage)
    static constexpr uint32_t warp_bytes(int mode) { return S * STAGE + (mode ? 2 * HALO_SLOT : 0) + (mode == 2 ? S * 128 : 0); }
    static constexpr uint32_t smem(int mode) { return TW_WARPS * (warp_bytes(mode) + 8 * S) + 128; }
};

// what pass_prepare_kernel saved for the in-place RGB pass (device memory owned by the caller of the launcher)
struct PassSide {
    const uint8_t *rows;            // [frame][band][2][row_pitch]: every band's first and last source row
    size_t row_pitch;
    const uint8_t *border;          // [frame][2][row_pitch]: what image rows 0 / height-1 become
    const uint8_t *entries_base;    // [frame][row][entries] x 16 bytes: (pixel left of strip boundary k, pixel right of it, 8 unused), RGB in bytes 0..2 / 4..6
    size_t entry_pitch;             // bytes per row of entries (multiple of 16)
    int entries;                    // nstrips + 1
    uint8_t *next_rows, *next_entries;      // the other set (written by this pass for the next one) or nullptr
};

struct TmaStencilParams {
    CUtensorMap map_full;     // box {RS/8, 3, 1} of uint64 elements (unused when !HAS_FULL)
    CUtensorMap map_half;     // box {HS/8, 3, 1}
    uint8_t *dst;
    size_t dst_pitch, dst_stride;
    const uint8_t *halo_above, *halo_below;   // row-band mode: source rows -1 / height (or nullptr)
    // in-place RGB output (MODE 2, see sobel_tma_kernel): what pass_prepare_kernel saved before anything is overwritten
    CUtensorMap map_side;     // {nstrips+1 entries of 16 bytes, height, frames}: the two pixels either side of every strip boundary + border values; box = 2 entries x 3 rows
    const uint8_t *side_rows; // [frame][band][first, last source row][side_row_pitch]
    size_t side_row_pitch;
    const uint8_t *border_rows;   // [frame][top, bottom][side_row_pitch]: what image rows 0 and height-1 become (scratch plane's border rows, replicated)
    // MODE 2 with more passes to come (ompimg_gray_sobel_passes_dev): this pass writes, next to its output, what the NEXT pass needs at the
    // places pass_prepare_kernel saved for this one (a second set of side buffers, same layout), so only the first pass needs a preparation
    uint8_t *next_rows;       // [frame][band][first, last source row][side_row_pitch], or nullptr
    uint8_t *next_entries;    // [frame][row][nstrips+1] x 16 bytes, or nullptr
    size_t entry_pitch;
    int width, height;
    int nbands, rows_per_band, nstrips;
    int fold;                 // the last strip is at most 256 pixels wide: one warp takes it for two bands
    int tasks_per_frame;
    long long ntasks;         // nframes * tasks_per_frame, one task per warp
    uint32_t row_bytes16;     // round16(width*CH)
};

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// source rows a band consumes: [rin0, rin1]; it produces rows rin0+1 .. rin1-1.  Band 0 takes one output row more than
// the others, so that with rows_per_band = 1 (mod 3) every band consumes a multiple of three source rows.
template <bool HALO>
__device__ __forceinline__ void band_rows(const TmaStencilParams &p, int band, int &rin0, int &rin1) {
    const int ya = band == 0 ? 0 : band * p.rows_per_band + 1, yb = min(p.height, (band + 1) * p.rows_per_band + 1);
    rin0 = max(ya - 1, HALO && p.halo_above ? -1 : 0);
    rin1 = band < p.nbands ? min(yb, HALO && p.halo_below ? p.height : p.height - 1) : rin0 - 1;      // a band that does not exist consumes nothing
}

template <int CH>
struct RowRaw {
    uint32_t w[4 * CH];
    uint32_t e0, e1;          // plane: the words left / right of the group; RGB(A): the word holding the strip-edge pixel
};
template <int CH>
__device__ __forceinline__ void load_row(uint32_t addr, int edge_off, RowRaw<CH> &r) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const uint4 v = lds128(addr + 16 * i);
        r.w[4 * i] = v.x; r.w[4 * i + 1] = v.y; r.w[4 * i + 2] = v.z; r.w[4 * i + 3] = v.w;
    }
    if (CH == 1) { r.e0 = lds32(addr - 4); r.e1 = lds32(addr + 16); }
    else r.e0 = lds32(addr + edge_off);
}

struct EdgeCtx {              // fused path: which lanes take a neighbour's luma from the halo block instead of a shuffle
    int edge_off;
    uint32_t edge_sh, mL, mR;
};

// P[k] = (v[2k], v[2k+1]), Q[k] = (v[2k-1], v[2k]) as f16x2 pairs 1024+v, from a plane row or fused from RGB(A)
template <int CH, int FMA_MODE>
__device__ __forceinline__ void make_pairs(const RowRaw<CH> &r, const EdgeCtx &ec, uint32_t (&P)[8], uint32_t (&Q)[9]) {
    if (CH == 1) {
        const uint32_t C64 = 0x64646464u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            P[2 * j] = prmt(r.w[j], C64, 0x4140);
            P[2 * j + 1] = prmt(r.w[j], C64, 0x4342);
            Q[2 * j + 1] = prmt(r.w[j], C64, 0x4241);
        }
        Q[0] = prmt(prmt(r.e0, r.w[0], 0x0043), C64, 0x4140);                // (lw.b3, w0.b0) + f16 exponents: two PRMT
#pragma unroll
        for (int j = 1; j < 4; ++j) Q[2 * j] = prmt(P[2 * j - 1], P[2 * j], 0x5432);
        Q[8] = prmt(prmt(r.w[3], r.e1, 0x0043), C64, 0x4140);
    } else {
        uint32_t f[16];
        const float MAGIC = 14942208.0f;      // 2^23 + 0x640000: result bits 0x4B6400ll
        if (CH == 3) luma16_rgb<FMA_MODE>(reinterpret_cast<const uint32_t(&)[12]>(r.w), f, MAGIC);
        else luma16_rgba<FMA_MODE>(reinterpret_cast<const uint32_t(&)[16]>(r.w), f, MAGIC);
        const uint32_t fE = luma_px<FMA_MODE>(r.e0 >> ec.edge_sh, MAGIC);
        uint32_t fL = __shfl_up_sync(0xffffffffu, f[15], 1);
        uint32_t fR = __shfl_down_sync(0xffffffffu, f[0], 1);
        fL = (fE & ec.mL) | (fL & ~ec.mL);
        fR = (fE & ec.mR) | (fR & ~ec.mR);
#pragma unroll
        for (int k = 0; k < 8; ++k) P[k] = prmt(f[2 * k], f[2 * k + 1], 0x6420);
        Q[0] = prmt(fL, f[0], 0x6420);
#pragma unroll
        for (int k = 1; k < 8; ++k) Q[k] = prmt(f[2 * k - 1], f[2 * k], 0x6420);
        Q[8] = prmt(f[15], fR, 0x6420);
    }
}

// The vertical recurrences rotate through three registers per pixel pair, so that every update is in place and
// the register roles are back where they started after three rows (no copies at the loop back-edge):
//   U[a] = Dh(r-2)+Dh(r-1), U[b] = Dh(r-1), U[c] free:  U[c] = Dh(r); U[b] += U[c]; gx = U[a]+U[b];  then a,b,c <- b,c,a
//   V[a] = Sh'(r-2), V[b] = Sh'(r-1), V[c] free:        V[c] = Sh'(r); gy = V[a]-V[c];               then a,b,c <- b,c,a
template <int RHO>
__device__ __forceinline__ void sobel_step(const uint32_t (&P)[8], const uint32_t (&Q)[9], uint32_t (&U)[3][8], uint32_t (&V)[3][8],
                                           uint32_t (&gx)[8], uint32_t (&gy)[8]) {
    constexpr int ra = RHO, rb = (RHO + 1) % 3, rc = (RHO + 2) % 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        U[rc][k] = hsub2(Q[k + 1], Q[k]);
        const uint32_t t1 = hfma2(P[k], 0x40004000u /*2,2*/, 0xEA00EA00u /*-3072,-3072*/);
        V[rc][k] = hadd2(hadd2(t1, Q[k]), Q[k + 1]);
        U[rb][k] = hadd2(U[rb][k], U[rc][k]);
        gx[k] = hadd2(U[ra][k], U[rb][k]);
        gy[k] = hsub2(V[ra][k], V[rc][k]);
    }
}

// RGBOUT epilogue (main_with_sobel.c:65-73): 16 edge bytes -> 16 RGB pixels, three 128-bit stores
__device__ __forceinline__ void store_rgb48(uint8_t *out, const uint4 &o) {
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
    uint32_t r[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        r[3 * q] = prmt(w[q], w[q], 0x1000);
        r[3 * q + 1] = prmt(w[q], w[q], 0x2211);
        r[3 * q + 2] = prmt(w[q], w[q], 0x3332);
    }
    uint4 *d = reinterpret_cast<uint4 *>(out);
    d[0] = make_uint4(r[0], r[1], r[2], r[3]);
    d[1] = make_uint4(r[4], r[5], r[6], r[7]);
    d[2] = make_uint4(r[8], r[9], r[10], r[11]);
}
// the group that straddles the end of a row whose width is not a multiple of 16: pixel by pixel (one lane per row)
__device__ __forceinline__ void store_rgb_partial(uint8_t *out, const uint4 &o, int npx) {
#pragma unroll 1
    for (int j = 0; j < npx; ++j) {
        const uint32_t w = j < 8 ? (j < 4 ? o.x : o.y) : (j < 12 ? o.z : o.w);
        const uint8_t v = (uint8_t)(w >> (8 * (j & 3)));
        out[3 * j] = v; out[3 * j + 1] = v; out[3 * j + 2] = v;
    }
}
__device__ __forceinline__ void set_byte16(uint4 &o, int i, uint32_t v) {       // byte i (0..15) of the 16-byte group := v
    const uint32_t sh = 8u * (uint32_t)(i & 3), m = ~(255u << sh), b = v << sh;
    if (i < 4) o.x = (o.x & m) | b;
    else if (i < 8) o.y = (o.y & m) | b;
    else if (i < 12) o.z = (o.z & m) | b;
    else o.w = (o.w & m) | b;
}

template <int N> struct IntC { static constexpr int value = N; };
template <bool B> struct BoolC { static constexpr bool value = B; };

template <int CH, int FMA_MODE, bool BIAS, int MODE>
__global__ void __launch_bounds__(TW_WARPS * 32, CH == 1 ? SOBEL_MINB_PLANE : SOBEL_MINB_RGB) sobel_tma_kernel(const __grid_constant__ TmaStencilParams p) {
    using G = TmaGeom<CH>;
    constexpr bool HALO = MODE == 1, RGBOUT = MODE == 2, SUBST = MODE != 0;
    static_assert(!RGBOUT || CH == 3, "in-place RGB output is built for 3 channels");
    constexpr int S = G::S;
    constexpr uint32_t RS = G::RS, HS = G::HS, STAGE = G::STAGE, HALFSTAGE = G::HALFSTAGE;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);           // warp-uniform by construction
    const long long task = (long long)blockIdx.x * TW_WARPS + warp;
    if (task >= p.ntasks) return;                       // warps are independent: no CTA-wide barrier below
    const int frame = (int)(task / p.tasks_per_frame);
    const int tf = (int)(task - (long long)frame * p.tasks_per_frame);
    // full strips first (strip-major inside a band, so that the warps of a CTA read one contiguous row piece),
    // then the folded tasks: the narrow last strip of bands 2i and 2i+1 in the two halves of one warp
    const int nfull = p.fold ? p.nstrips - 1 : p.nstrips;
    const bool folded = tf >= p.nbands * nfull;
    const int strip = folded ? p.nstrips - 1 : tf % nfull;
    const int band_a = folded ? 2 * (tf - p.nbands * nfull) : tf / nfull;
    const bool split = !G::HAS_FULL || folded;          // two half-strip boxes per stage
    const int hb = split ? lane >> 4 : 0;               // which half of the warp
    const int lc = split ? lane & 15 : lane;            // 16-pixel group of the lane inside its (half) strip
    const int xs = strip * 512;

    int a0, a1, b0, b1;                                 // source rows of the two halves
    band_rows<HALO>(p, band_a, a0, a1);
    if (folded) band_rows<HALO>(p, band_a + 1, b0, b1);
    else { b0 = a0; b1 = split ? a1 : a0 - 1; }         // RGBA: both halves walk the same band; full layout: no second half
    const int nin_a = a1 - a0 + 1, nin_b = b1 - b0 + 1;
    const int nin_max = max(nin_a, nin_b);
    if (nin_max < 3) return;
    const int ntrips = (nin_max + 2) / 3;
    const bool has_b = split && nin_b >= 3;

    const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
    const uint32_t ring0 = base + (uint32_t)warp * G::warp_bytes(MODE);
    const uint32_t bar0 = base + TW_WARPS * G::warp_bytes(MODE) + (uint32_t)warp * (8 * S);
    const uint32_t hslot0 = ring0 + S * STAGE;          // SUBST: the two spare row slots (the band's first / last source row)
    const uint32_t sbox0 = hslot0 + 2 * G::HALO_SLOT;   // RGBOUT: one 128-byte side box per stage (3 rows x 16 bytes used)

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();

    // tensor coordinates of the boxes: x in 8-byte elements (the left halo block of strip 0 lies left of the row: zero-filled)
    const int cxa = (xs * CH - 16) / 8;
    const int cxb = folded ? cxa : cxa + 256 * CH / 8;
    // row-band mode: which source rows of this band come from the neighbours, and where they sit in the trip sequence
    // (in-place RGB output: EVERY band takes its first and last source row from the copies saved by pass_prepare_kernel)
    const bool use_above = RGBOUT || (HALO && p.halo_above != nullptr && a0 < 0);
    const bool use_below = RGBOUT || (HALO && p.halo_below != nullptr && a1 >= p.height);
    const int below_i = RGBOUT ? nin_a - 1 : p.height - a0;      // index of that last row in the band's row sequence
    const int below_t = below_i / 3, below_rho = below_i - 3 * below_t;
    const uint8_t *row_first = RGBOUT ? p.side_rows + ((size_t)frame * p.nbands + band_a) * 2 * p.side_row_pitch : p.halo_above;
    const uint8_t *row_last = RGBOUT ? row_first + p.side_row_pitch : p.halo_below;

    // elected lane: one 1-D bulk copy per (half-)strip piece of a neighbour's row into a spare row slot
    auto peer_row_bytes = [&]() -> uint32_t {
        uint32_t n = 0;
        const int parts = split ? 2 : 1, wpx = split ? 256 : 512;
        for (int q = 0; q < parts; ++q) {
            const uint32_t xq = (uint32_t)(xs + q * 256);
            const uint32_t lo = xq == 0 ? 0u : xq * CH - 16u, hi = min(p.row_bytes16, (xq + wpx) * CH + 16u);
            if (hi > lo) n += hi - lo;
        }
        return n;
    };
    auto peer_row_copy = [&](uint32_t slot, const uint8_t *row, uint32_t bar) {
        const int parts = split ? 2 : 1, wpx = split ? 256 : 512;
        for (int q = 0; q < parts; ++q) {
            const uint32_t xq = (uint32_t)(xs + q * 256);
            const uint32_t lo = xq == 0 ? 0u : xq * CH - 16u, hi = min(p.row_bytes16, (xq + wpx) * CH + 16u);
            if (hi > lo) bulk_g2s(slot + q * HS + (xq == 0 ? 16u : 0u), row + lo, hi - lo, bar);
        }
    };
    // elected lane: request trip tt (source rows a0+3tt .. +2, and b0+3tt .. for the second half) into stage s
    auto request = [&](int s, int tt) {
        const uint32_t bar = bar0 + 8 * s, dst = ring0 + (uint32_t)s * STAGE;
        uint32_t tx = split ? (has_b ? 6u * HS : 3u * HS) : 3u * RS;
        if (SUBST) {
            if (use_above && tt == 0) tx += peer_row_bytes();
            if (use_below && tt == below_t) tx += peer_row_bytes();
        }
        if (RGBOUT) tx += 96u;
        mbar_expect_tx(bar, tx);
        if (G::HAS_FULL && !split) {
            tma_load_3d(dst, &p.map_full, cxa, a0 + 3 * tt, frame, bar);
        } else {
            tma_load_3d(dst, &p.map_half, cxa, a0 + 3 * tt, frame, bar);
            if (has_b) tma_load_3d(dst + HALFSTAGE, &p.map_half, cxb, b0 + 3 * tt, frame, bar);
        }
        if (SUBST) {
            if (use_above && tt == 0) peer_row_copy(hslot0, row_first, bar);
            if (use_below && tt == below_t) peer_row_copy(hslot0 + G::HALO_SLOT, row_last, bar);
        }
        if (RGBOUT) tma_load_3d(sbox0 + (uint32_t)s * 128u, &p.map_side, 2 * strip, a0 + 3 * tt, frame, bar);      // entries strip, strip+1 of three rows (a box must START on a 16-byte boundary of global memory: hence 16-byte entries)
    };
    if (elect_one_sync()) {
        for (int t = 0; t < S && t < ntrips; ++t) request(t, t);
    }

    const int rin0 = hb ? b0 : a0;
    const int nin = hb ? nin_b : nin_a;                 // rows this lane's half really has
    const int x0 = xs + (split && !folded ? 256 * hb : 0) + 16 * lc;
    // How this lane stores its 16 output bytes: 0 nothing, 1 one 16-byte store, 2 the group holds x == 0, 3 it holds
    // x == width-1 (and possibly bytes past the row), 4 both.  Border groups must leave dst's border bytes as they are
    // (sobel.c:11-12 never writes them): a few narrow stores, in the few warps that hold such a group.  (Reading the group
    // back and merging was tried: the dependent load stalls those warps for a DRAM latency per row -- 0.75 -> 0.49.)
    // (in-place RGB output writes the border pixels too -- they become the scratch plane's border values,
    // main_with_sobel.c:65-73 -- so there every lane inside the image stores its whole group)
    int mode = 0;
    if (x0 < p.width && nin >= 3) {
        const bool l = x0 == 0, r = x0 + 16 > p.width - 1;
        mode = l ? (r ? 4 : 2) : (r ? 3 : 1);
    }
    const int plast = min(16, p.width - 1 - x0);
    const int npx = min(16, p.width - x0);              // RGBOUT: pixels of the group inside the image
    const bool edge_warp = __any_sync(0xffffffffu, mode >= 2);       // warp-uniform
    const uint32_t rs = split ? HS : RS;                                  // bytes between the rows of a stage
    const uint32_t my = ring0 + (uint32_t)hb * HALFSTAGE + 16u + (uint32_t)lc * (16u * CH);
    const uint32_t my_halo = hslot0 + (uint32_t)hb * HS + 16u + (uint32_t)lc * (16u * CH);
    // fused path: the first lane of a (half) strip takes its left neighbour's luma from the halo block, the last lane its right one
    EdgeCtx ec;
    {
        const bool first = lc == 0, last = lc == (split ? 15 : 31);
        ec.edge_off = first ? -4 : 16 * CH;
        ec.edge_sh = first && !RGBOUT ? 8u * (4u - CH) : 0u;
        ec.mL = first ? 0xffffffffu : 0u;
        ec.mR = last ? 0xffffffffu : 0u;
    }
    const size_t dpitch = p.dst_pitch;
    uint8_t *out = p.dst + (size_t)frame * p.dst_stride + ((ptrdiff_t)rin0 - 1) * (ptrdiff_t)p.dst_pitch + (size_t)x0 * (RGBOUT ? 3 : 1);   // stores go one row above the source row being consumed
    const uint32_t side_off = ec.mL ? 0u : 20u;         // RGBOUT: the strip-edge pixel of this lane in a side-box row (first lane: left one, last lane: right one)
    uint32_t side_prev = 0;                             // ... and the previous row's, whose low byte is that row's border value in the outermost lanes
    // next pass's side data: lane 0 writes the pixel right of strip boundary `strip` (its first pixel), lane 31 the pixel left of boundary
    // `strip + 1` (its last pixel, when the image reaches that far), of every row this warp writes
    const bool ent_lane = RGBOUT && p.next_entries != nullptr && (lane == 0 || (lane == 31 && xs + 511 < p.width));
    uint8_t *ent = nullptr;
    if (RGBOUT && ent_lane)
        ent = p.next_entries + ((ptrdiff_t)frame * p.height + rin0 - 1) * (ptrdiff_t)p.entry_pitch + (lane == 0 ? strip * 16 + 4 : (strip + 1) * 16);      // like `out`: one row above the source row being consumed
    const size_t ent_pitch = RGBOUT ? p.entry_pitch : 0;
    float half = BIAS ? 0.5f : 0.0f;
    if (BIAS) asm volatile("" : "+f"(half));

    uint32_t U[3][8], V[3][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) U[0][k] = U[1][k] = U[2][k] = V[0][k] = V[1][k] = V[2][k] = 0u;

    int s = 0;
    uint32_t ph = 0;

    // one source row: load, (after the stage's last row: release the stage by requesting its refill,) arithmetic, store
    auto row = [&](auto rho_c, auto checked_c, auto edge_c, uint32_t addr, int t) {
        constexpr int RHO = decltype(rho_c)::value;
        constexpr bool CHECKED = decltype(checked_c)::value;
        constexpr bool EDGE = decltype(edge_c)::value;           // this warp holds a group on the image border (warp-uniform, fixed for the task)
        if (SUBST && CHECKED) {
            if (use_above && t == 0 && RHO == 0) addr = my_halo;
            if (use_below && t == below_t && RHO == below_rho) addr = my_halo + G::HALO_SLOT;
        }
        RowRaw<CH> raw;
        load_row<CH>(addr, ec.edge_off, raw);
        if (RGBOUT) raw.e0 = lds32(sbox0 + (uint32_t)s * 128u + (uint32_t)RHO * 32u + side_off);      // the halo blocks of the stage may be overwritten already
        if (RHO == 2) {
            // the refill is an async-proxy WRITE of the stage the generic-proxy loads above READ: the proxy fence orders each
            // lane's loads before it, the warp barrier collects the lanes (conv_tma.cuh has the failure this prevents)
            fence_async_smem();
            __syncwarp();                                   // every lane holds the stage's three rows
            if (elect_one_sync() && t + S < ntrips) request(s, t + S);
        }
        uint32_t P[8], Q[9], gx[8], gy[8];
        make_pairs<CH, FMA_MODE>(raw, ec, P, Q);
        sobel_step<RHO>(P, Q, U, V, gx, gy);
        const bool emit = !CHECKED || (3 * t + RHO >= 2 && 3 * t + RHO < nin);       // the first two rows only prime the row state
        if (emit) {
            uint4 o = sobel_mag16(gx, gy, half);                    // once per row, whatever the lanes' store modes
            if (!RGBOUT) {
                if (mode == 1) *reinterpret_cast<uint4 *>(out) = o;
                if (EDGE) {
                    if (mode >= 2) store_edge16(out, o, mode, plast);
                }
            } else {
                if (EDGE) {
                    // the image border takes the scratch plane's border values of the OUTPUT row (the previous source row's side entries)
                    const uint32_t br = __shfl_sync(0xffffffffu, side_prev, 31) & 255u;
                    if (mode == 2 || mode == 4) o.x = (o.x & 0xffffff00u) | (side_prev & 255u);
                    if (mode >= 3) set_byte16(o, plast, br);
                }
                if (mode) {
                    if (npx == 16) store_rgb48(out, o);
                    else store_rgb_partial(out, o, npx);
                }
                if (ent_lane) *reinterpret_cast<uint32_t *>(ent) = (lane == 0 ? (o.x & 255u) : (o.w >> 24)) * 0x010101u;
                if (CHECKED && p.next_rows != nullptr && mode) {
                    // this band's first output row is the previous band's last source row, its last output row the next band's first one
                    const int i = 3 * t + RHO;
                    if (i == 2 && band_a >= 1) {
                        uint8_t *q = p.next_rows + (((size_t)frame * p.nbands + (band_a - 1)) * 2 + 1) * p.side_row_pitch + (size_t)x0 * 3;
                        if (npx == 16) store_rgb48(q, o); else store_rgb_partial(q, o, npx);
                    }
                    if (i == nin - 1 && band_a + 1 < p.nbands) {
                        uint8_t *q = p.next_rows + (((size_t)frame * p.nbands + (band_a + 1)) * 2 + 0) * p.side_row_pitch + (size_t)x0 * 3;
                        if (npx == 16) store_rgb48(q, o); else store_rgb_partial(q, o, npx);
                    }
                }
            }
        }
        if (RGBOUT) { side_prev = raw.e0; ent += ent_pitch; }
        out += dpitch;
    };
    auto trip = [&](auto checked_c, auto edge_c, int t) {
        mbar_wait(bar0 + 8 * s, ph);
        const uint32_t a = my + (uint32_t)s * STAGE;
        row(IntC<0>{}, checked_c, edge_c, a, t);
        row(IntC<1>{}, checked_c, edge_c, a + rs, t);
        row(IntC<2>{}, checked_c, edge_c, a + 2 * rs, t);
        if (++s == S) { s = 0; ph ^= 1; }
    };

    if (RGBOUT) {
        // image rows 0 and height-1 become the scratch plane's border rows: copied by the strips of the first / last band
        // (nobody reads those rows of the image any more: bands take them from the saved copies)
        const uint8_t *brow = p.border_rows + (size_t)frame * 2 * p.side_row_pitch + (size_t)x0 * 3;
        uint8_t *irow = p.dst + (size_t)frame * p.dst_stride + (size_t)x0 * 3;
        for (int which = 0; which < 2; ++which) {
            if (which == 0 ? a0 != 0 : a1 != p.height - 1) continue;
            const uint8_t *sr = brow + (size_t)which * p.side_row_pitch;
            uint8_t *dr = irow + (size_t)(which ? p.height - 1 : 0) * p.dst_pitch;
            // ... and, for the next pass, the same row as band 0's first / the last band's last source row, and its strip-boundary entries
            uint8_t *nr = p.next_rows ? p.next_rows + (((size_t)frame * p.nbands + band_a) * 2 + which) * p.side_row_pitch + (size_t)x0 * 3 : nullptr;
            if (npx == 16) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const uint4 v = reinterpret_cast<const uint4 *>(sr)[q];
                    reinterpret_cast<uint4 *>(dr)[q] = v;
                    if (nr) reinterpret_cast<uint4 *>(nr)[q] = v;
                }
            } else {
                for (int j = 0; j < 3 * npx; ++j) { dr[j] = sr[j]; if (nr) nr[j] = sr[j]; }
            }
            if (ent_lane) {
                const uint8_t *px = lane == 0 ? sr : sr + 45;
                uint8_t *e = p.next_entries + ((size_t)frame * p.height + (size_t)(which ? p.height - 1 : 0)) * p.entry_pitch +
                             (lane == 0 ? (size_t)strip * 16 + 4 : (size_t)(strip + 1) * 16);
                *reinterpret_cast<uint32_t *>(e) = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
            }
        }
    }

    // trips [1, tsteady) lie entirely inside every storing lane's band and hold no substituted row
    int tsteady = (has_b ? min(nin_a, nin_b) : nin_a) / 3;
    if (SUBST && use_below) tsteady = min(tsteady, below_t);
    // warps without a border group (six of a 4K row's eight strips) run a steady body with no border-store code at all
    // (not the in-place RGB mode: a third copy of its larger body costs more in instruction fetch than it saves -- 119 -> 98 Gpx/s)
    trip(BoolC<true>{}, BoolC<true>{}, 0);
    int t = 1;
    if (RGBOUT || edge_warp) for (; t < tsteady; ++t) trip(BoolC<false>{}, BoolC<true>{}, t);
    else for (; t < tsteady; ++t) trip(BoolC<false>{}, BoolC<RGBOUT>{}, t);
    for (; t < ntrips; ++t) trip(BoolC<true>{}, BoolC<true>{}, t);
}

// cuTensorMapEncodeTiled comes from the driver through the runtime (the library links the static runtime only)
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        else
            cudaGetLastError();
    }
    return fn;
}
// {row bytes / 8, height, frames} of uint64 elements, box {box_bytes / 8, 3, 1}
static cudaError_t make_row_map(CUtensorMap *m, const uint8_t *base, uint32_t row_bytes16, size_t pitch, int height, int nframes, size_t stride,
                                uint32_t box_bytes) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[3] = {row_bytes16 / 8, (cuuint64_t)height, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {pitch, nframes > 1 ? stride : pitch * (size_t)height};
    const cuuint32_t box[3] = {box_bytes / 8, 3, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// band / strip decomposition of a launch (shared by the kernel launch and by the in-place pass's preparation)
struct TmaDecomp { int nstrips, rows_per_band, nbands; };
static TmaDecomp tma_decomp(int width, int height, int nframes, int ch, int sm_count) {
    TmaDecomp d;
    d.nstrips = (width + 511) / 512;
    // Band height (1 mod 3: each band then consumes a multiple of three source rows), chosen to waste least: every band
    // re-reads two rows and primes on two rows (2/rows extra), and a launch of T tasks on the resident warp slots
    // (24 per SM for planes, 16 for RGB(A)) runs in ceil(T/slots) waves of which the last is only partly full.
    // Launches of fewer than ~5 waves do not consider bands taller than 64 rows: in more, shorter waves the warps of an SM drift
    // apart and keep the memory system evenly loaded, which is worth more than the 2/rows saved -- 64 4K frames: fused 1381 Gpx/s
    // with the 112-row bands the estimate alone picks, 1431-1460 with 40-64 rows; planes 2479 -> 2523, 2000x3000 planes 2376 -> 2478.
    // Long launches keep the tall bands (256 frames, 32768^2: 2 % better with 94-130 rows; tools/prof.py, OMPIMG_SOBEL_BAND).
    const double slots = (double)sm_count * TW_WARPS * (ch == 1 ? SOBEL_MINB_PLANE : SOBEL_MINB_RGB);
    const int cmax = (double)nframes * d.nstrips * ((height - 1 + 129) / 130) < 5.0 * slots ? 64 : 130;
    int rpb = cmax;
    double best = -1.0;
    for (int c = cmax; c >= 16; c -= 3) {
        const double tasks = (double)nframes * d.nstrips * ((height - 1 + c - 1) / c);
        const double waves = tasks / slots, full = waves <= 1.0 ? 1.0 : (double)(long long)(waves + 0.999999);
        const double eff = waves / full * (double)c / (double)(c + 2);
        if (eff > best * 1.005) { best = eff; rpb = c; }        // ties go to the taller band
    }
    { static int env = -1; if (env < 0) { const char *e = getenv("OMPIMG_SOBEL_BAND"); env = e ? atoi(e) : 0; } if (env > 0) rpb = env; }
    d.rows_per_band = rpb;
    d.nbands = max(1, (height - 1 + rpb - 1) / rpb);
    return d;
}

template <int CH, int FMA_MODE, bool BIAS, int MODE>
static cudaError_t launch_sobel_tma_h(const StencilArgs &a, int sm_count, cudaStream_t st, const PassSide *side = nullptr) {
    using G = TmaGeom<CH>;
    constexpr bool HALO = MODE == 1;
    TmaStencilParams p{};
    p.dst = a.dst; p.dst_pitch = a.dst_pitch; p.dst_stride = a.dst_stride;
    p.width = a.width; p.height = a.height;
    p.halo_above = a.halo_above; p.halo_below = a.halo_below;
    const TmaDecomp dc = tma_decomp(a.width, a.height, a.nframes, CH, sm_count);
    p.nstrips = dc.nstrips; p.rows_per_band = dc.rows_per_band; p.nbands = dc.nbands;
    p.row_bytes16 = (uint32_t)(((size_t)a.width * CH + 15) & ~(size_t)15);
    // a last strip of at most 256 pixels (3840 = 7.5 strips) would leave half a warp idle: fold two bands into one warp
    { static int env = -1; if (env < 0) { const char *e = getenv("OMPIMG_SOBEL_FOLD"); env = e ? atoi(e) : 1; }
      p.fold = env && p.nbands >= 2 && a.width - (p.nstrips - 1) * 512 <= 256 && MODE == 0; }
    p.tasks_per_frame = p.fold ? p.nbands * (p.nstrips - 1) + (p.nbands + 1) / 2 : p.nbands * p.nstrips;
    p.ntasks = (long long)a.nframes * p.tasks_per_frame;
    const long long ctas = (p.ntasks + TW_WARPS - 1) / TW_WARPS;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    cudaError_t e;
    if (G::HAS_FULL) {
        e = make_row_map(&p.map_full, a.src, p.row_bytes16, a.src_pitch, a.height, a.nframes, a.src_stride, G::RS);
        if (e != cudaSuccess) return e;
    }
    e = make_row_map(&p.map_half, a.src, p.row_bytes16, a.src_pitch, a.height, a.nframes, a.src_stride, G::HS);
    if (e != cudaSuccess) return e;
    if (MODE == 2) {
        if (!side) return cudaErrorInvalidValue;
        p.side_rows = side->rows; p.side_row_pitch = side->row_pitch; p.border_rows = side->border;
        p.next_rows = side->next_rows; p.next_entries = side->next_entries; p.entry_pitch = side->entry_pitch;
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return cudaErrorNotSupported;
        const cuuint64_t dims[3] = {(cuuint64_t)side->entries * 2, (cuuint64_t)a.height, (cuuint64_t)a.nframes};      // uint64 elements
        const cuuint64_t strides[2] = {side->entry_pitch, side->entry_pitch * (size_t)a.height};
        const cuuint32_t box[3] = {4, 3, 1}, estr[3] = {1, 1, 1};
        if (enc(&p.map_side, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t *>(side->entries_base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    constexpr int smem = (int)G::smem(MODE);
    static unsigned long long attr_mask = 0;
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(attr_mask >> dev & 1ULL)) {
        e = cudaFuncSetAttribute(sobel_tma_kernel<CH, FMA_MODE, BIAS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_mask |= 1ULL << dev;
    }
    count_launch();
    sobel_tma_kernel<CH, FMA_MODE, BIAS, MODE><<<(unsigned)ctas, TW_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <int CH, int FMA_MODE, bool BIAS>
static cudaError_t launch_sobel_tma_t(const StencilArgs &a, int sm_count, cudaStream_t st) {
    return (a.halo_above || a.halo_below) ? launch_sobel_tma_h<CH, FMA_MODE, BIAS, 1>(a, sm_count, st)
                                          : launch_sobel_tma_h<CH, FMA_MODE, BIAS, 0>(a, sm_count, st);
}

// ---------------------------------------------------------------------------------------------
// The in-place pass (a5, ompimg_gray_sobel_pass_dev with 3 channels): pass_prepare_kernel, then sobel_tma_kernel<3,...,2>.
// ---------------------------------------------------------------------------------------------
struct PrepParams {
    const uint8_t *img; size_t pitch, stride;
    int w, h, nframes, nbands, rpb, nstrips;
    uint32_t row_bytes16;
    const uint8_t *edge; size_t epitch, estride;         // the driver's scratch plane (its border is what the image border becomes) or nullptr = zero
    uint8_t *rows; size_t row_pitch;                     // out: [frame][band][2][row_pitch]
    uint8_t *border; /* out: [frame][2][row_pitch] */
    uint8_t *entries; size_t entry_pitch;                // out: [frame][row][nstrips+1] x 16 bytes
    uint8_t *entries2;                                   // out (multi-pass runs): the other set's constant halves -- the row's border values in entry 0's left and the last entry's right pixel -- or nullptr
};
// Reads the image only; everything the in-place kernel could find overwritten is copied out first (see sobel_tma_kernel).
// Three ranges of CTAs, no 64-bit divisions: [0, nA) copy every band's first and last source row (one thread per 16-byte chunk),
// [nA, nA+nC) build the replicated border rows, the rest write the strip-boundary entries (one thread per entry, two aligned
// 32-bit loads: the pixel left of boundary k ends at byte 1536k, the pixel right of it starts there).
__global__ void __launch_bounds__(256) pass_prepare_kernel(const PrepParams p) {
    const uint32_t chunks = p.row_bytes16 / 16;
    const uint32_t cpr = (chunks + 255u) / 256u;                                // CTAs per copied row
    const uint32_t nA = (uint32_t)p.nframes * (uint32_t)p.nbands * 2u * cpr, nC = (uint32_t)p.nframes * 2u * cpr;
    const uint32_t blk = blockIdx.x;
    if (blk < nA) {
        const uint32_t q = blk / cpr, c = (blk - q * cpr) * 256u + threadIdx.x;  // q = (frame * nbands + band) * 2 + which
        if (c >= chunks) return;
        const uint32_t which = q & 1u, fb = q >> 1, f = fb / (uint32_t)p.nbands, bnd = fb - f * (uint32_t)p.nbands;
        const int a0 = (int)bnd * p.rpb, yb = min(p.h, ((int)bnd + 1) * p.rpb + 1), a1 = min(yb, p.h - 1);
        if (a1 - a0 + 1 < 3) return;                                            // a band without output rows is never consumed
        const int r = which ? a1 : a0;
        const uint4 v = *reinterpret_cast<const uint4 *>(p.img + (size_t)f * p.stride + (size_t)r * p.pitch + (size_t)c * 16);
        *reinterpret_cast<uint4 *>(p.rows + (size_t)q * p.row_pitch + (size_t)c * 16) = v;
    } else if (blk < nA + nC) {
        const uint32_t j = blk - nA, q = j / cpr, c = (j - q * cpr) * 256u + threadIdx.x;       // q = frame * 2 + which
        if (c >= chunks) return;
        const uint32_t which = q & 1u, f = q >> 1;
        uint32_t o[4] = {0, 0, 0, 0};
        if (p.edge) {
            const uint8_t *er = p.edge + (size_t)f * p.estride + (size_t)(which ? p.h - 1 : 0) * p.epitch;
            const uint32_t lim = (uint32_t)p.w * 3u;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t byte = c * 16u + k;
                const uint32_t v = byte < lim ? er[byte / 3u] : 0u;
                o[k >> 2] |= v << (8 * (k & 3));
            }
        }
        *reinterpret_cast<uint4 *>(p.border + (size_t)q * p.row_pitch + (size_t)c * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        const uint32_t ne = (uint32_t)p.nstrips + 1u;
        const unsigned long long item = (unsigned long long)(blk - nA - nC) * 256u + threadIdx.x;     // (frame * h + y) * ne + k
        if (item >= (unsigned long long)p.nframes * (unsigned long long)p.h * ne) return;
        const uint32_t fy = (uint32_t)(item / ne), k = (uint32_t)(item - (unsigned long long)fy * ne);      // nframes * h < 2^32 (checked by the launcher)
        const uint32_t f = fy / (uint32_t)p.h, y = fy - f * (uint32_t)p.h;
        const uint8_t *row = p.img + (size_t)f * p.stride + (size_t)y * p.pitch;
        const uint8_t *er = p.edge ? p.edge + (size_t)f * p.estride + (size_t)y * p.epitch : nullptr;
        const uint32_t xb = k * 512u;
        uint32_t L, R;
        if (k == 0) L = (er ? (uint32_t)er[0] : 0u) * 0x010101u;
        else if (xb - 1u < (uint32_t)p.w) L = *reinterpret_cast<const uint32_t *>(row + (size_t)xb * 3 - 4) >> 8;      // bytes 1536k-3 .. 1536k-1
        else L = 0u;                                                            // no pixel there (the last strip is ragged): never looked at
        if (xb < (uint32_t)p.w) R = *reinterpret_cast<const uint32_t *>(row + (size_t)xb * 3) & 0xffffffu;
        else R = (er ? (uint32_t)er[p.w - 1] : 0u) * 0x010101u;
        *reinterpret_cast<uint4 *>(p.entries + (size_t)fy * p.entry_pitch + (size_t)k * 16) = make_uint4(L, R, 0u, 0u);
        if (p.entries2) {
            uint32_t *e2 = reinterpret_cast<uint32_t *>(p.entries2 + (size_t)fy * p.entry_pitch + (size_t)k * 16);
            if (k == 0) e2[0] = L;
            if (k == (uint32_t)p.nstrips) e2[1] = R;
        }
    }
}

bool pass_inplace_supported(const uint8_t *img, size_t pitch, size_t stride, int w, int h, int ch, int nframes) {
    return ch == 3 && w >= 3 && h >= 3 && ((uintptr_t)img & 15) == 0 && (pitch & 15) == 0 && (nframes == 1 || (stride & 15) == 0) &&
           pitch >= (((size_t)w * 3 + 15) & ~(size_t)15);
}
PassScratch pass_inplace_scratch(int w, int h, int nframes, int sm_count) {
    const TmaDecomp d = tma_decomp(w, h, nframes, 3, sm_count);
    PassScratch s;
    s.row_pitch = ((size_t)w * 3 + 15) & ~(size_t)15;
    s.rows_bytes = (size_t)nframes * d.nbands * 2 * s.row_pitch;
    s.border_bytes = (size_t)nframes * 2 * s.row_pitch;
    s.entry_pitch = (size_t)(d.nstrips + 1) * 16;
    s.entries_bytes = (size_t)nframes * h * s.entry_pitch;
    return s;
}
template <int FMA_MODE, bool BIAS>
static cudaError_t launch_pass_inplace_t(const StencilArgs &a, const PassSide &side, int sm_count, cudaStream_t st) {
    return launch_sobel_tma_h<3, FMA_MODE, BIAS, 2>(a, sm_count, st, &side);
}
// `passes` iterations of main_with_sobel.c:51-74 in place on `img` (3 channels).  scratch = rows | border | entries [| rows | entries for
// passes > 1] as sized by pass_inplace_scratch: the preparation kernel fills the first set from the image; every pass but the last
// writes the other set for its successor, so a run of n passes is 1 + n launches.
cudaError_t launch_gray_sobel_pass_inplace(uint8_t *img, size_t pitch, size_t stride, int w, int h, int nframes, const uint8_t *edge, size_t epitch,
                                           size_t estride, int fma_mode, int no_bias, int passes, uint8_t *scratch, int sm_count, cudaStream_t st) {
    if (passes < 1) return cudaSuccess;
    const TmaDecomp d = tma_decomp(w, h, nframes, 3, sm_count);
    const PassScratch sz = pass_inplace_scratch(w, h, nframes, sm_count);
    uint8_t *rows[2] = {scratch, scratch + sz.total()};
    uint8_t *border = scratch + sz.rows_bytes;
    uint8_t *entries[2] = {scratch + sz.rows_bytes + sz.border_bytes, scratch + sz.total() + sz.rows_bytes};
    PrepParams pp{};
    pp.img = img; pp.pitch = pitch; pp.stride = stride; pp.w = w; pp.h = h; pp.nframes = nframes;
    pp.nbands = d.nbands; pp.rpb = d.rows_per_band; pp.nstrips = d.nstrips;
    pp.row_bytes16 = (uint32_t)sz.row_pitch;
    pp.edge = edge; pp.epitch = epitch; pp.estride = estride;
    pp.rows = rows[0]; pp.row_pitch = sz.row_pitch;
    pp.border = border;
    pp.entries = entries[0]; pp.entry_pitch = sz.entry_pitch;
    pp.entries2 = passes > 1 ? entries[1] : nullptr;
    const long long cpr = ((long long)(sz.row_pitch / 16) + 255) / 256;
    const long long ctas = (long long)nframes * (d.nbands + 1) * 2 * cpr + ((long long)nframes * h * (d.nstrips + 1) + 255) / 256;
    if (ctas > 0x7fffffffLL || (long long)nframes * h > 0xffffffffLL) return cudaErrorInvalidValue;
    count_launch();
    pass_prepare_kernel<<<(int)ctas, 256, 0, st>>>(pp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    StencilArgs a{};
    a.src = img; a.dst = img; a.src_pitch = pitch; a.dst_pitch = pitch; a.src_stride = stride; a.dst_stride = stride;
    a.width = w; a.height = h; a.channels = 3; a.nframes = nframes; a.fma_mode = fma_mode; a.no_bias = no_bias;
    for (int ps = 0; ps < passes; ++ps) {
        const int cur = ps & 1, more = ps + 1 < passes;
        PassSide side{};
        side.rows = rows[cur]; side.row_pitch = sz.row_pitch; side.border = border;
        side.entries_base = entries[cur]; side.entry_pitch = sz.entry_pitch; side.entries = d.nstrips + 1;
        side.next_rows = more ? rows[cur ^ 1] : nullptr; side.next_entries = more ? entries[cur ^ 1] : nullptr;
        if (no_bias) e = fma_mode ? launch_pass_inplace_t<1, false>(a, side, sm_count, st) : launch_pass_inplace_t<0, false>(a, side, sm_count, st);
        else e = fma_mode ? launch_pass_inplace_t<1, true>(a, side, sm_count, st) : launch_pass_inplace_t<0, true>(a, side, sm_count, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// generic fallback (any alignment / pitch): one thread per output pixel
// ---------------------------------------------------------------------------------------------
template <int CH, int FMA_MODE>
__device__ __forceinline__ int px_value(const uint8_t *row, int x) {
    if (CH == 1) return row[x];
    return (int)luma1<FMA_MODE>(row[x * CH], row[x * CH + 1], row[x * CH + 2]);
}
__device__ __forceinline__ uint32_t isqrt_clamp255(uint32_t m) {
    if (m >= 65025u) return 255u;
    uint32_t r = (uint32_t)sqrtf((float)m);      // correctly rounded sqrt of an exact float; fix the edges anyway
    while (r * r > m) --r;
    while ((r + 1) * (r + 1) <= m) ++r;
    return r;
}
template <int CH, int FMA_MODE>
__global__ void __launch_bounds__(256) sobel_generic_kernel(const StencilArgs a) {
    const long long iw = a.width - 2;
    const int ylo = (a.halo_above ? 0 : 1), yhi = (a.halo_below ? a.height - 1 : a.height - 2);   // inclusive output rows
    const long long ih = yhi - ylo + 1;
    if (iw <= 0 || ih <= 0) return;
    const long long total = (long long)a.nframes * ih * iw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / (ih * iw), rem = i - f * ih * iw;
        const int y = ylo + (int)(rem / iw), x = 1 + (int)(rem % iw);
        const uint8_t *base = a.src + f * a.src_stride;
        const uint8_t *ra = (y - 1 < 0) ? a.halo_above : base + (size_t)(y - 1) * a.src_pitch;
        const uint8_t *rb = base + (size_t)y * a.src_pitch;
        const uint8_t *rc = (y + 1 >= a.height) ? a.halo_below : base + (size_t)(y + 1) * a.src_pitch;
        const int a0 = px_value<CH, FMA_MODE>(ra, x - 1), a1 = px_value<CH, FMA_MODE>(ra, x), a2 = px_value<CH, FMA_MODE>(ra, x + 1);
        const int b0 = px_value<CH, FMA_MODE>(rb, x - 1), b2 = px_value<CH, FMA_MODE>(rb, x + 1);
        const int c0 = px_value<CH, FMA_MODE>(rc, x - 1), c1 = px_value<CH, FMA_MODE>(rc, x), c2 = px_value<CH, FMA_MODE>(rc, x + 1);
        const int gx = -a0 - 2 * b0 - c0 + a2 + 2 * b2 + c2;
        const int gy = a0 + 2 * a1 + a2 - c0 - 2 * c1 - c2;
        a.dst[f * a.dst_stride + (size_t)y * a.dst_pitch + x] = (uint8_t)isqrt_clamp255((uint32_t)(gx * gx + gy * gy));
    }
}

template <int CH, int FMA_MODE, bool BIAS>
static cudaError_t launch_sobel_t(const StencilArgs &a, int sm_count, cudaStream_t st) {
    if (a.width < 3 || a.nframes < 1) return cudaSuccess;
    const int out_rows = a.height - (a.halo_above ? 0 : 1) - (a.halo_below ? 0 : 1);
    if (out_rows < 1) return cudaSuccess;
    const bool fast = (((uintptr_t)a.src | (uintptr_t)a.dst | (uintptr_t)a.halo_above | (uintptr_t)a.halo_below) & 15) == 0 &&
                      ((a.src_pitch | a.dst_pitch) & 15) == 0 && (a.nframes == 1 || ((a.src_stride | a.dst_stride) & 15) == 0) &&
                      a.src_pitch >= (((size_t)a.width * CH + 15) & ~(size_t)15);
    if (!fast) {
        const long long total = (long long)a.nframes * out_rows * (a.width - 2);
        const int grid = (int)max(1LL, min((total + 255) / 256, (long long)sm_count * 16));
        count_launch();
        sobel_generic_kernel<CH, FMA_MODE><<<grid, 256, 0, st>>>(a);
        return cudaGetLastError();
    }
    if (a.height + (a.halo_above ? 1 : 0) + (a.halo_below ? 1 : 0) < 3) return cudaSuccess;
    return launch_sobel_tma_t<CH, FMA_MODE, BIAS>(a, sm_count, st);
}

cudaError_t launch_sobel(const StencilArgs &a, int sm_count, cudaStream_t st) {
    if (a.no_bias) {
        switch (a.channels) {
            case 1: return launch_sobel_t<1, 0, false>(a, sm_count, st);
            case 3: return a.fma_mode ? launch_sobel_t<3, 1, false>(a, sm_count, st) : launch_sobel_t<3, 0, false>(a, sm_count, st);
            case 4: return a.fma_mode ? launch_sobel_t<4, 1, false>(a, sm_count, st) : launch_sobel_t<4, 0, false>(a, sm_count, st);
        }
        return cudaErrorInvalidValue;
    }
    switch (a.channels) {
        case 1: return launch_sobel_t<1, 0, true>(a, sm_count, st);
        case 3: return a.fma_mode ? launch_sobel_t<3, 1, true>(a, sm_count, st) : launch_sobel_t<3, 0, true>(a, sm_count, st);
        case 4: return a.fma_mode ? launch_sobel_t<4, 1, true>(a, sm_count, st) : launch_sobel_t<4, 0, true>(a, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// self tests of the magnitude path
// ---------------------------------------------------------------------------------------------
// out[m] = sat_u8(trunc(sqrt.approx(m + 0.5)))  -- must equal min(255, isqrt(m)) for every m
__global__ void isqrt_selftest_kernel(uint8_t *out, unsigned int count, float bias) {
    for (unsigned int m = blockIdx.x * blockDim.x + threadIdx.x; m < count; m += gridDim.x * blockDim.x)
        out[m] = (uint8_t)(pack_sat_u8(0.f, sqrt_approx((float)m + bias), 0u) & 255u);
}
cudaError_t launch_isqrt_selftest(uint8_t *out_dev, unsigned int count, float bias, cudaStream_t st) {
    count_launch();
    isqrt_selftest_kernel<<<1024, 256, 0, st>>>(out_dev, count, bias);
    return cudaGetLastError();
}

}  // namespace ompimg
