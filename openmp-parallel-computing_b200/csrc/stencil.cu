// stencil.cu -- 3x3 Sobel (a2: monolithic/src/sobel.c:6-25) and the fused RGB -> luma -> Sobel path
// (a3..a5: monolithic/src/main_with_sobel.c:51-74), as one row-streaming kernel.
//
// Decomposition: a CTA owns a band of rows of one frame and a column strip of up to NW*512 pixels.
// A producer warp streams the band's source rows into a ring of shared-memory stages with 1-D bulk
// asynchronous copies (TMA engine, completion on an mbarrier).  NW consumer warps each own 512
// columns (16 pixels per lane) and walk down the rows keeping the two previous rows' horizontal
// partial results in registers (f16x2 pairs), so every source byte is read from HBM once, touched
// in shared memory once, and the luma plane of the fused path never exists in memory.  The only
// neighbour exchange is one luma value per warp edge per row (shuffles inside a warp, 8 bytes of
// shared memory between warps).  Filter taps are immediates folded into the f16x2 arithmetic.
//
// Row-band mode (ompimg_bands_*): the rows just above/below the band live on a peer GPU; the
// producer warp fetches them with ordinary global loads through the NVLink peer mapping.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ompimg {

constexpr int ST_MAX_NW = 8;      // consumer warps per CTA

struct StencilParams {
    const uint8_t *src;
    uint8_t *dst;
    size_t src_pitch, dst_pitch, src_stride, dst_stride;
    const uint8_t *halo_above, *halo_below;
    int width, height;
    int nbands, rows_per_band;
    int nw;             // consumer warps
    int stages;
    uint32_t stage_bytes;   // three row slots
    uint32_t row_slot;      // bytes reserved per row in a stage (16-byte halo blocks included)
    uint32_t row_bytes16;   // round16(width*bpp)
    int relaxed;            // consumers sleep (instead of re-probing) when their stage is not refilled yet
};

template <int CH>
__device__ __forceinline__ void load_group_words(uint32_t addr, uint32_t (&w)[4 * CH]) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const uint4 v = lds128(addr + 16 * i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
}

template <int CH, int FMA_MODE, bool BIAS>
__global__ void __launch_bounds__((ST_MAX_NW + 1) * 32, CH == 1 ? 3 : 2) sobel_rows_kernel(const StencilParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int BPP = CH;                      // bytes per source pixel
    const int NW = p.nw, S = p.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + S * p.stage_bytes;     // S full barriers, then S empty barriers
    const uint32_t empty0 = full0 + 8 * S;

    const int frame = blockIdx.x / p.nbands, band = blockIdx.x - frame * p.nbands;
    const int SW = NW * 512;
    const int xs = blockIdx.y * SW;
    const uint8_t *src = p.src + (size_t)frame * p.src_stride;
    uint8_t *dst = p.dst + (size_t)frame * p.dst_stride;

    // band 0 takes rows_per_band+1 output rows, every other band rows_per_band: with rows_per_band = 1 (mod 3)
    // each band then consumes a multiple of three source rows (the row loop below runs three rows per trip)
    const int ya = band == 0 ? 0 : band * p.rows_per_band + 1, yb = min(p.height, (band + 1) * p.rows_per_band + 1);
    const int rin0 = max(ya - 1, p.halo_above ? -1 : 0);
    const int rin1 = min(yb, p.halo_below ? p.height : p.height - 1);
    const int nin = rin1 - rin0 + 1;              // source rows this CTA consumes; outputs rows rin0+1 .. rin1-1

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NW); }
        mbar_fence_init();
    }
    __syncthreads();
    if (nin < 3) return;

    if (warp == NW) {
        // ------------------------------------------------------------------ producer warp
        const uint32_t lo = xs == 0 ? 0u : (uint32_t)xs * BPP - 16u;
        const uint32_t hi = min(p.row_bytes16, (uint32_t)(xs + SW) * BPP + 16u);
        const uint32_t bytes = hi - lo;
        const uint32_t dst_off = xs == 0 ? 16u : 0u;
        int s = 0;
        uint32_t ph = 0;
        const int ntrips = (nin + 2) / 3;         // a stage holds three consecutive rows: one wait / one release per trip
        for (int t = 0; t < ntrips; ++t) {
            if (t >= S) mbar_wait_relaxed(empty0 + 8 * s, ph ^ 1);
            const uint32_t d0 = smem0 + s * p.stage_bytes + dst_off;
            uint32_t tx = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int r = min(rin0 + 3 * t + j, rin1);
                if (r < 0 || r >= p.height) {
                    // halo row owned by a neighbouring band (peer device memory over NVLink): plain loads
                    const uint8_t *row = (r < 0 ? p.halo_above : p.halo_below) + lo;
                    for (uint32_t q = lane * 16; q < bytes; q += 32 * 16) sts128(d0 + j * p.row_slot + q, *reinterpret_cast<const uint4 *>(row + q));
                } else {
                    tx += bytes;
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(full0 + 8 * s, tx);             // also releases the stores above to the consumers
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int r = min(rin0 + 3 * t + j, rin1);
                    if (r >= 0 && r < p.height) bulk_g2s(d0 + j * p.row_slot, src + (size_t)r * p.src_pitch + lo, bytes, full0 + 8 * s);
                }
            }
            if (++s == S) { s = 0; ph ^= 1; }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    // Everything loop-invariant is computed once and pinned in a register with keep(): ptxas
    // otherwise rematerialises these values (S2R + LDC + integer math) inside the row loop, and
    // this kernel is bound by issue slots, not by registers.
    const int c = warp * 32 + lane;               // 16-pixel group index inside the strip
    const int x0 = xs + 16 * c;
    // how this lane stores its 16 output bytes: 0 nothing, 1 one 16-byte store, 2 group holds x == 0,
    // 3 group holds x == width-1 (and possibly bytes past the row), 4 both (very narrow images)
    int mode = 0;
    if (x0 < p.width) {
        const bool l = x0 == 0, r = x0 + 16 > p.width - 1;
        mode = l ? (r ? 4 : 2) : (r ? 3 : 1);
    }
    mode = (int)keep((uint32_t)mode);
    const int plast = min(16, p.width - 1 - x0);  // valid bytes are [x0 == 0, plast)
    const uint32_t my = keep(smem0 + 16u + (uint32_t)c * (16u * BPP));
    const uint32_t stage_bytes = keep(p.stage_bytes);
    const uint32_t row_slot = keep(p.row_slot);
    const uint32_t ring_bytes = keep((uint32_t)S * p.stage_bytes);
    const uint32_t eoff = keep(8u * (uint32_t)S);
    const uint32_t fb0 = keep(full0);
    const uint32_t lane0 = keep(lane == 0 ? 1u : 0u);
    // warp-edge lumas of the fused path: lane 0 needs the luma of the pixel left of its group, lane 31 of the
    // pixel right of its group.  Both pixels are in the stage (neighbour warp's bytes, or the 16-byte
    // halo blocks at the strip ends), so every warp recomputes its own two edge lumas: one extra pixel
    // of luma arithmetic per row instead of a CTA-wide barrier, and the warps stay independent.
    const int edge_off = lane == 0 ? -4 : 16 * BPP;                         // word that holds the edge pixel
    const uint32_t edge_sh = keep(lane == 0 ? 8u * (4u - BPP) : 0u);         // its bytes start here
    const uint32_t mL = keep(lane == 0 ? 0xffffffffu : 0u);
    const uint32_t mR = keep(lane == 31 ? 0xffffffffu : 0u);
    const size_t dpitch = keep64(p.dst_pitch);
    uint8_t *out = dst + ((ptrdiff_t)rin0 - 1) * (ptrdiff_t)p.dst_pitch + x0;     // stores go one row above the source row being consumed
    float half = BIAS ? 0.5f : 0.0f;
    if (BIAS) asm volatile("" : "+f"(half));

    // row state per pixel pair, in two three-register rotations (see the row loop)
    uint32_t U[3][8], V[3][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) U[0][k] = U[1][k] = U[2][k] = V[0][k] = V[1][k] = V[2][k] = 0u;

    uint32_t so = 0;                 // byte offset of the current stage in the ring
    uint32_t fb = fb0;               // its full barrier (the empty barrier sits 8*S bytes above)
    uint32_t ph = 0;

    // Rows are consumed three per trip.  Per pixel pair the vertical recurrences rotate through three
    // registers each, so that every update is in place and the register roles are back where they
    // started after three rows (no copies at the loop back-edge):
    //   U[a] = Dh(r-2)+Dh(r-1), U[b] = Dh(r-1), U[c] free:  U[c] = Dh(r); U[b] += U[c]; gx = U[a]+U[b];  then a,b,c <- b,c,a
    //   V[a] = Sh'(r-2), V[b] = Sh'(r-1), V[c] free:        V[c] = Sh'(r); gy = V[a]-V[c];               then a,b,c <- b,c,a
    // A row count that is not a multiple of three is padded with dummy rows whose results are not
    // stored (the producer re-sends the last row for them).
    // (Fetching the next row's words before the sqrt/pack/store tail was tried and measured slower:
    // the extra live registers cost more than the exposed shared-memory latency.)
    for (int i = 0; i < nin; i += 3) {
#pragma unroll
      for (int rho = 0; rho < 3; ++rho) {
        constexpr int dummy_ = 0; (void)dummy_;
        const int ra = rho, rb = (rho + 1) % 3, rc = (rho + 2) % 3;
        const bool emit = i + rho >= 2 && i + rho < nin;        // the first two rows only prime the row state
        const uint32_t stage = my + so + (uint32_t)rho * row_slot;
        // (a warp that finds its stage not yet refilled is ahead of the slowest warp of the CTA -- the ring is only
        // refilled when every warp has released a stage; sleeping there instead of re-probing measured the same)
        if (rho == 0) { if (p.relaxed) mbar_wait_relaxed(fb, ph); else mbar_wait(fb, ph); }
        uint32_t P[8], Q[9];
        if (CH == 1) {
            const uint32_t C64 = 0x64646464u;
            const uint4 v = lds128(stage);
            const uint32_t lw = lds32(stage - 4), rw = lds32(stage + 16);
            __syncwarp();
            if (rho == 2 && lane0) mbar_arrive(fb + eoff);   // all three rows of the stage are in registers / consumed
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                P[2 * j] = prmt(w[j], C64, 0x4140);
                P[2 * j + 1] = prmt(w[j], C64, 0x4342);
                Q[2 * j + 1] = prmt(w[j], C64, 0x4241);
            }
            Q[0] = prmt(prmt(lw, w[0], 0x0043), C64, 0x4140);                // (lw.b3, w0.b0) + f16 exponents: two PRMT
#pragma unroll
            for (int j = 1; j < 4; ++j) Q[2 * j] = prmt(P[2 * j - 1], P[2 * j], 0x5432);
            Q[8] = prmt(prmt(w[3], rw, 0x0043), C64, 0x4140);
        } else {
            uint32_t w[4 * CH];
            load_group_words<CH>(stage, w);
            const uint32_t ew = lds32(stage + edge_off) >> edge_sh;          // r,g,b of the edge pixel in bytes 0..2
            __syncwarp();
            if (rho == 2 && lane0) mbar_arrive(fb + eoff);   // all three rows of the stage are in registers / consumed
            uint32_t f[16];
            const float MAGIC = 14942208.0f;      // 2^23 + 0x640000: result bits 0x4B6400ll
            if (CH == 3) luma16_rgb<FMA_MODE>(reinterpret_cast<const uint32_t(&)[12]>(w), f, MAGIC);
            else luma16_rgba<FMA_MODE>(reinterpret_cast<const uint32_t(&)[16]>(w), f, MAGIC);
            const uint32_t fE = luma_px<FMA_MODE>(ew, MAGIC);
            uint32_t fL = __shfl_up_sync(0xffffffffu, f[15], 1);
            uint32_t fR = __shfl_down_sync(0xffffffffu, f[0], 1);
            fL = (fE & mL) | (fL & ~mL);
            fR = (fE & mR) | (fR & ~mR);
#pragma unroll
            for (int k = 0; k < 8; ++k) P[k] = prmt(f[2 * k], f[2 * k + 1], 0x6420);
            Q[0] = prmt(fL, f[0], 0x6420);
#pragma unroll
            for (int k = 1; k < 8; ++k) Q[k] = prmt(f[2 * k - 1], f[2 * k], 0x6420);
            Q[8] = prmt(f[15], fR, 0x6420);
        }
        uint32_t gx[8], gy[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            U[rc][k] = hsub2(Q[k + 1], Q[k]);
            const uint32_t t1 = hfma2(P[k], 0x40004000u /*2,2*/, 0xEA00EA00u /*-3072,-3072*/);
            V[rc][k] = hadd2(hadd2(t1, Q[k]), Q[k + 1]);
            U[rb][k] = hadd2(U[rb][k], U[rc][k]);
            gx[k] = hadd2(U[ra][k], U[rb][k]);
            gy[k] = hsub2(V[ra][k], V[rc][k]);
        }
        if (emit && mode) {
            const uint4 o = sobel_mag16(gx, gy, half);
            if (mode == 1) *reinterpret_cast<uint4 *>(out) = o;
            else store_edge16(out, o, mode, plast);
        }
        out += dpitch;
        if (rho == 2) {
            so += stage_bytes;
            fb += 8;
            if (so == ring_bytes) { so = 0; fb = fb0; ph ^= 1; }
        }
      }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous variant (the default): every warp is its own pipeline.
//
// A warp owns a 512-pixel column strip of a band of rows and runs a private ring of stages (three
// rows each) with a private mbarrier per stage; an elected lane issues the strip's bulk copies
// (512*CH + 2x16 bytes per row) and the whole warp consumes them -- there is no producer warp, no
// empty barrier and no coupling between the warps of a CTA, so a warp never waits for a slower
// neighbour before its next rows can be requested (in sobel_rows_kernel the ring is refilled only
// when all eight warps have released a stage, and the consumers measured ~20 % of their time in
// that wait).  The CTA is 8 warps = 256 threads, so three CTAs per SM get 80 registers instead of
// 72 (288 threads), which removes the spills of the plane kernel.  Same arithmetic, same results.
// ---------------------------------------------------------------------------------------------
struct WarpStencilParams {
    const uint8_t *src;
    uint8_t *dst;
    size_t src_pitch, dst_pitch, src_stride, dst_stride;
    int width, height;
    int nbands, rows_per_band, nstrips;
    const uint8_t *halo_above, *halo_below;   // row-band mode: the source rows -1 / height, in a peer GPU's memory (or nullptr)
    int fold;               // the last strip is at most 256 pixels wide: one warp takes it for TWO bands (16 lanes each)
    int tasks_per_frame;
    long long ntasks;       // nframes * tasks_per_frame, one task per warp
    uint32_t row_bytes16;   // round16(width*CH)
};
constexpr int WS_WARPS = 8;
template <int CH> constexpr int ws_stages() { return CH == 1 ? 4 : 2; }
template <int CH> constexpr int ws_half_slot() { return 32 + 256 * CH; }          // half a strip and its two 16-byte halo blocks
template <int CH> constexpr int ws_row_slot() { return 2 * ws_half_slot<CH>(); }
template <int CH> constexpr int ws_smem() { return WS_WARPS * (ws_stages<CH>() * 3 * ws_row_slot<CH>() + 8 * ws_stages<CH>() + 64) + 128; }

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// rows a band consumes / produces (band 0 takes one output row more, see sobel_rows_kernel)
template <bool HALO>
__device__ __forceinline__ void ws_band_rows(const WarpStencilParams &p, int band, int &rin0, int &rin1) {
    const int ya = band == 0 ? 0 : band * p.rows_per_band + 1, yb = min(p.height, (band + 1) * p.rows_per_band + 1);
    rin0 = max(ya - 1, HALO && p.halo_above ? -1 : 0);
    rin1 = band < p.nbands ? min(yb, HALO && p.halo_below ? p.height : p.height - 1) : rin0 - 1;      // a band that does not exist consumes nothing
}

// HALO: row-band mode, rows -1 / height come from p.halo_above / p.halo_below (a separate instantiation: the extra
// tests in the request path cost the ordinary launches 5-8 % when compiled in)
template <int CH, int FMA_MODE, bool BIAS, bool HALO>
__global__ void __launch_bounds__(WS_WARPS * 32, CH == 1 ? 3 : 2) sobel_warp_kernel(const __grid_constant__ WarpStencilParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int BPP = CH, S = ws_stages<CH>();
    constexpr uint32_t HALF_SLOT = ws_half_slot<CH>(), ROW_SLOT = ws_row_slot<CH>(), STAGE = 3 * ROW_SLOT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * WS_WARPS + warp;
    if (task >= p.ntasks) return;                       // warps are independent: no CTA-wide barrier below
    const long long frame = task / p.tasks_per_frame;
    const int tf = (int)(task - frame * p.tasks_per_frame);
    // full strips first (strip-major inside a band, so that the warps of a CTA read one contiguous row piece),
    // then the folded tasks: the narrow last strip of bands 2i and 2i+1 in the two halves of one warp
    const int nfull = p.fold ? p.nstrips - 1 : p.nstrips;
    const bool folded = tf >= p.nbands * nfull;
    const int strip = folded ? p.nstrips - 1 : tf % nfull;
    const int band_a = folded ? 2 * (tf - p.nbands * nfull) : tf / nfull;
    const int hb = folded ? lane >> 4 : 0;              // which half of the warp (folded tasks)
    const int lc = folded ? lane & 15 : lane;           // 16-pixel group of the lane inside its (half) strip
    const uint32_t ring0 = (smem_u32(smem) + 127u & ~127u) + (uint32_t)warp * (S * STAGE);
    const uint32_t bar0 = (smem_u32(smem) + 127u & ~127u) + WS_WARPS * (S * STAGE) + (uint32_t)warp * (8 * S);
    const uint32_t scr0 = (smem_u32(smem) + 127u & ~127u) + WS_WARPS * (S * STAGE + 8 * S) + (uint32_t)warp * 64u;   // see request()

    const int xs = strip * 512;
    const uint8_t *src = p.src + (size_t)frame * p.src_stride;
    uint8_t *dst = p.dst + (size_t)frame * p.dst_stride;
    int a0, a1, b0, b1;                                 // source rows of the two halves' bands (half B only in folded tasks)
    ws_band_rows<HALO>(p, band_a, a0, a1);
    ws_band_rows<HALO>(p, folded ? band_a + 1 : p.nbands, b0, b1);
    const int nin_a = a1 - a0 + 1, nin_b = b1 - b0 + 1;
    const int nin_max = max(nin_a, nin_b);
    if (nin_max < 3) return;
    const int ntrips = (nin_max + 2) / 3;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();

    // What the elected lane needs to request rows lives in the warp's scratch words, not in registers of
    // all 32 lanes (the kernel is register-bound): address and index of the next source row of each half,
    // their last rows, bytes per row, stage offset of the first byte.
    // The strip's bytes of a row are [lo, hi), 16-byte halo blocks included where they exist.
    if (lane == 0) {
        const uint32_t lo = xs == 0 ? 0u : (uint32_t)xs * BPP - 16u;
        const uint32_t hi = min(p.row_bytes16, (uint32_t)(xs + (folded ? 256 : 512)) * BPP + 16u);
        const u64 ga = (u64)((long long)(uintptr_t)(src + lo) + (long long)a0 * (long long)p.src_pitch);     // a0 is -1 with a halo row above: never dereferenced
        const u64 gb = (u64)(uintptr_t)(src + (size_t)max(b0, 0) * p.src_pitch + lo);
        sts128(scr0, make_uint4((uint32_t)ga, (uint32_t)(ga >> 32), (uint32_t)gb, (uint32_t)(gb >> 32)));
        sts128(scr0 + 16, make_uint4((uint32_t)a0, (uint32_t)b0, (uint32_t)a1, (uint32_t)b1));
        sts32(scr0 + 32, hi - lo); sts32(scr0 + 36, (xs == 0 ? 16u : 0u) | (nin_b >= 3 ? 256u : 0u));
        sts32(scr0 + 40, lo);
    }
    __syncwarp();
    auto request = [&](int s) {       // elected lane: the next three rows of each half into stage s (past a band's last row: that row again)
        const uint4 c0 = lds128(scr0), c1 = lds128(scr0 + 16);
        const uint32_t bytes = lds32(scr0 + 32), fl = lds32(scr0 + 36);
        u64 ga = pack64(c0.x, c0.y), gb = pack64(c0.z, c0.w);
        int ra_next = (int)c1.x, rb_next = (int)c1.y;
        const int ra1 = (int)c1.z, rb1 = (int)c1.w;
        const bool lb = (fl & 256u) != 0;
        const uint32_t d0 = ring0 + s * STAGE + (fl & 255u), bar = bar0 + 8 * s;
        // rows -1 and height of a row band live in a peer GPU's memory
        const bool peer_a = HALO && ((p.halo_above && ra_next < 0) || (p.halo_below && ra_next + 2 >= p.height && ra1 >= p.height));
        if (HALO && peer_a && !lb) {
            const uint32_t lo = lds32(scr0 + 40);
            // the peer rows come by the same bulk copies, straight over NVLink from the neighbour's allocation
            mbar_expect_tx(bar, 3u * bytes);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int r = min(ra_next + j, ra1);
                const void *row = (r >= 0 && r < p.height) ? reinterpret_cast<const void *>((uintptr_t)ga)
                                                           : static_cast<const void *>((r < 0 ? p.halo_above : p.halo_below) + lo);
                bulk_g2s(d0 + j * ROW_SLOT, row, bytes, bar);
                if (ra_next + j < ra1) ga += p.src_pitch;
            }
            ra_next = min(ra_next + 3, ra1);
        } else if (!lb && ra_next + 3 <= ra1) {
            mbar_expect_tx(bar, 3u * bytes);
            // the common case (no second half, not at the band's end): three consecutive rows
#pragma unroll
            for (int j = 0; j < 3; ++j) bulk_g2s(d0 + j * ROW_SLOT, reinterpret_cast<const void *>((uintptr_t)(ga + j * p.src_pitch)), bytes, bar);
            ga += 3 * p.src_pitch;
            ra_next += 3;
        } else {
            mbar_expect_tx(bar, (lb ? 6u : 3u) * bytes);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                bulk_g2s(d0 + j * ROW_SLOT, reinterpret_cast<const void *>((uintptr_t)ga), bytes, bar);
                if (ra_next < ra1) { ga += p.src_pitch; ++ra_next; }
                if (lb) {
                    bulk_g2s(d0 + j * ROW_SLOT + HALF_SLOT, reinterpret_cast<const void *>((uintptr_t)gb), bytes, bar);
                    if (rb_next < rb1) { gb += p.src_pitch; ++rb_next; }
                }
            }
        }
        uint32_t g0, g1, g2, g3;
        unpack64(ga, g0, g1); unpack64(gb, g2, g3);
        sts128(scr0, make_uint4(g0, g1, g2, g3));
        sts32(scr0 + 16, (uint32_t)ra_next); sts32(scr0 + 20, (uint32_t)rb_next);
    };
    if (elect_one_sync()) {
        for (int t = 0; t < S && t < ntrips; ++t) request(t);
    }

    const int rin0 = hb ? b0 : a0;
    const int nin = hb ? nin_b : nin_a;                 // rows this lane's half really has (the loop runs nin_max rows)
    const int x0 = xs + 16 * lc;
    int mode = 0;         // 0 nothing to store, 1 one 16-byte store, 2 group holds x == 0, 3 group holds x == width-1, 4 both
    if (x0 < p.width && nin >= 3) {
        const bool l = x0 == 0, r = x0 + 16 > p.width - 1;
        mode = l ? (r ? 4 : 2) : (r ? 3 : 1);
    }
    const int plast = min(16, p.width - 1 - x0);
    const uint32_t my = ring0 + (uint32_t)hb * HALF_SLOT + 16u + (uint32_t)lc * (16u * BPP);
    // fused path: the first lane of a (half) strip takes its left neighbour's luma from the halo block, lane 31 its right one
    const bool first = lc == 0;
    const int edge_off = first ? -4 : 16 * BPP;
    const uint32_t edge_sh = first ? 8u * (4u - BPP) : 0u;
    const uint32_t mL = first ? 0xffffffffu : 0u, mR = lane == 31 ? 0xffffffffu : 0u;
    const size_t dpitch = p.dst_pitch;
    uint8_t *out = dst + ((ptrdiff_t)rin0 - 1) * (ptrdiff_t)p.dst_pitch + x0;     // stores go one row above the source row being consumed
    float half = BIAS ? 0.5f : 0.0f;
    if (BIAS) asm volatile("" : "+f"(half));

    uint32_t U[3][8], V[3][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) U[0][k] = U[1][k] = U[2][k] = V[0][k] = V[1][k] = V[2][k] = 0u;

    int s = 0, t = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nin_max; i += 3, ++t) {
        mbar_wait(bar0 + 8 * s, ph);
#pragma unroll
        for (int rho = 0; rho < 3; ++rho) {
            const int ra = rho, rb = (rho + 1) % 3, rc = (rho + 2) % 3;
            const bool emit = i + rho >= 2 && i + rho < nin;
            const uint32_t stage = my + (uint32_t)s * STAGE + (uint32_t)rho * ROW_SLOT;
            uint32_t P[8], Q[9];
            if (CH == 1) {
                const uint32_t C64 = 0x64646464u;
                const uint4 v = lds128(stage);
                const uint32_t lw = lds32(stage - 4), rw = lds32(stage + 16);
                if (rho == 2) {
                    __syncwarp();                                   // every lane holds the stage's three rows
                    if (elect_one_sync() && t + S < ntrips) request(s);
                }
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    P[2 * j] = prmt(w[j], C64, 0x4140);
                    P[2 * j + 1] = prmt(w[j], C64, 0x4342);
                    Q[2 * j + 1] = prmt(w[j], C64, 0x4241);
                }
                Q[0] = prmt(prmt(lw, w[0], 0x0043), C64, 0x4140);                // (lw.b3, w0.b0) + f16 exponents: two PRMT
#pragma unroll
                for (int j = 1; j < 4; ++j) Q[2 * j] = prmt(P[2 * j - 1], P[2 * j], 0x5432);
                Q[8] = prmt(prmt(w[3], rw, 0x0043), C64, 0x4140);
            } else {
                uint32_t w[4 * CH];
                load_group_words<CH>(stage, w);
                const uint32_t ew = lds32(stage + edge_off) >> edge_sh;
                if (rho == 2) {
                    __syncwarp();
                    if (elect_one_sync() && t + S < ntrips) request(s);
                }
                uint32_t f[16];
                const float MAGIC = 14942208.0f;
                if (CH == 3) luma16_rgb<FMA_MODE>(reinterpret_cast<const uint32_t(&)[12]>(w), f, MAGIC);
                else luma16_rgba<FMA_MODE>(reinterpret_cast<const uint32_t(&)[16]>(w), f, MAGIC);
                const uint32_t fE = luma_px<FMA_MODE>(ew, MAGIC);
                uint32_t fL = __shfl_up_sync(0xffffffffu, f[15], 1);
                uint32_t fR = __shfl_down_sync(0xffffffffu, f[0], 1);
                fL = (fE & mL) | (fL & ~mL);
                fR = (fE & mR) | (fR & ~mR);
#pragma unroll
                for (int k = 0; k < 8; ++k) P[k] = prmt(f[2 * k], f[2 * k + 1], 0x6420);
                Q[0] = prmt(fL, f[0], 0x6420);
#pragma unroll
                for (int k = 1; k < 8; ++k) Q[k] = prmt(f[2 * k - 1], f[2 * k], 0x6420);
                Q[8] = prmt(f[15], fR, 0x6420);
            }
            uint32_t gx[8], gy[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                U[rc][k] = hsub2(Q[k + 1], Q[k]);
                const uint32_t t1 = hfma2(P[k], 0x40004000u /*2,2*/, 0xEA00EA00u /*-3072,-3072*/);
                V[rc][k] = hadd2(hadd2(t1, Q[k]), Q[k + 1]);
                U[rb][k] = hadd2(U[rb][k], U[rc][k]);
                gx[k] = hadd2(U[ra][k], U[rb][k]);
                gy[k] = hsub2(V[ra][k], V[rc][k]);
            }
            if (emit && mode) {
                const uint4 o = sobel_mag16(gx, gy, half);
                if (mode == 1) *reinterpret_cast<uint4 *>(out) = o;
                else store_edge16(out, o, mode, plast);
            }
            out += dpitch;
        }
        if (++s == S) { s = 0; ph ^= 1; }
    }
}

template <int CH, int FMA_MODE, bool BIAS, bool HALO>
static cudaError_t launch_sobel_warp_h(const StencilArgs &a, int sm_count, cudaStream_t st) {
    WarpStencilParams p{};
    p.src = a.src; p.dst = a.dst; p.src_pitch = a.src_pitch; p.dst_pitch = a.dst_pitch;
    p.src_stride = a.src_stride; p.dst_stride = a.dst_stride;
    p.width = a.width; p.height = a.height;
    p.halo_above = a.halo_above; p.halo_below = a.halo_below;
    p.nstrips = (a.width + 511) / 512;
    p.row_bytes16 = (uint32_t)(((size_t)a.width * CH + 15) & ~(size_t)15);
    // bands of ~130 rows (1 mod 3: each band then consumes a multiple of three source rows); shorter when the launch is small
    // Pick the band height that wastes least: every band re-reads two rows and primes on two rows (2/rows extra), and
    // a launch of T tasks on S resident warp slots (24 per SM for planes, 16 for RGB(A)) runs in ceil(T/S) waves of
    // which the last is only partly full.
    int rpb = 130;
    {
        const double slots = (double)sm_count * (CH == 1 ? 24 : 16);
        double best = -1.0;
        for (int c = 130; c >= 16; c -= 3) {                                    // all 1 (mod 3)
            const double tasks = (double)a.nframes * p.nstrips * ((a.height - 1 + c - 1) / c);
            const double waves = tasks / slots, full = waves <= 1.0 ? 1.0 : (double)(long long)(waves + 0.999999);
            const double eff = waves / full * (double)c / (double)(c + 2);
            if (eff > best * 1.005) { best = eff; rpb = c; }        // ties go to the taller band
        }
    }
    { static int env = -1; if (env < 0) { const char *e = getenv("OMPIMG_SOBEL_BAND"); env = e ? atoi(e) : 0; } if (env > 0) rpb = env; }
    p.rows_per_band = rpb;
    p.nbands = max(1, (a.height - 1 + rpb - 1) / rpb);
    // a last strip of at most 256 pixels (3840 = 7.5 strips) would leave half a warp idle: fold two bands into one warp
    { static int env = -1; if (env < 0) { const char *e = getenv("OMPIMG_SOBEL_FOLD"); env = e ? atoi(e) : 1; }
      p.fold = env && p.nbands >= 2 && a.width - (p.nstrips - 1) * 512 <= 256 && !a.halo_above && !a.halo_below; }
    p.tasks_per_frame = p.fold ? p.nbands * (p.nstrips - 1) + (p.nbands + 1) / 2 : p.nbands * p.nstrips;
    p.ntasks = (long long)a.nframes * p.tasks_per_frame;
    const long long ctas = (p.ntasks + WS_WARPS - 1) / WS_WARPS;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    constexpr int smem = ws_smem<CH>();
    static unsigned long long attr_mask = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(attr_mask >> dev & 1ULL)) {
        e = cudaFuncSetAttribute(sobel_warp_kernel<CH, FMA_MODE, BIAS, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_mask |= 1ULL << dev;
    }
    count_launch();
    sobel_warp_kernel<CH, FMA_MODE, BIAS, HALO><<<(unsigned)ctas, WS_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <int CH, int FMA_MODE, bool BIAS>
static cudaError_t launch_sobel_warp_t(const StencilArgs &a, int sm_count, cudaStream_t st) {
    return (a.halo_above || a.halo_below) ? launch_sobel_warp_h<CH, FMA_MODE, BIAS, true>(a, sm_count, st)
                                          : launch_sobel_warp_h<CH, FMA_MODE, BIAS, false>(a, sm_count, st);
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
    {
        // warp-autonomous pipeline whenever no halo rows come from peers (OMPIMG_SOBEL_WARP=0: the row-ring kernel, for A/B runs)
        static int warp_mode = -1;
        if (warp_mode < 0) { const char *e = getenv("OMPIMG_SOBEL_WARP"); warp_mode = e ? atoi(e) : 1; }
        const bool use_warp = warp_mode != 0 && a.height + (a.halo_above ? 1 : 0) + (a.halo_below ? 1 : 0) >= 3;
        if (use_warp) return launch_sobel_warp_t<CH, FMA_MODE, BIAS>(a, sm_count, st);
    }
    StencilParams p{};
    p.src = a.src; p.dst = a.dst; p.src_pitch = a.src_pitch; p.dst_pitch = a.dst_pitch;
    p.src_stride = a.src_stride; p.dst_stride = a.dst_stride;
    p.halo_above = a.halo_above; p.halo_below = a.halo_below;
    p.width = a.width; p.height = a.height;
    { static int relaxed = -1; if (relaxed < 0) { const char *e = getenv("OMPIMG_RELAXED_WAIT"); relaxed = e ? atoi(e) : 0; } p.relaxed = relaxed; }
    p.nw = min(ST_MAX_NW, (a.width + 511) / 512);
    const int nstrips = (a.width + p.nw * 512 - 1) / (p.nw * 512);
    p.row_bytes16 = (uint32_t)(((size_t)a.width * CH + 15) & ~(size_t)15);
    p.row_slot = (uint32_t)((32 + p.nw * 512 * CH + 127) & ~127);
    p.stage_bytes = 3 * p.row_slot;
    // occupancy target: ~24-32 resident warps per SM
    const int occ = CH == 1 ? max(1, min(4, 30 / (p.nw + 1))) : max(1, min(2, 30 / (p.nw + 1)));   // CTAs per SM the registers allow
    const int budget = (227 * 1024) / occ - 1024 - 768;
    p.stages = max(2, min(5, budget / (int)p.stage_bytes));
    const int smem = p.stages * (int)p.stage_bytes + 16 * p.stages + 64;
    // bands: ~128 rows when there is plenty of work, fewer rows (>=16) when the launch is small
    const long long want_ctas = (long long)sm_count * occ * 4;
    int rpb = 130;                                    // 1 (mod 3), see the kernel
    while (rpb > 16 && (long long)a.nframes * nstrips * ((a.height - 1 + rpb - 1) / rpb) < want_ctas) rpb = rpb == 130 ? 64 : rpb == 64 ? 34 : 16;
    p.rows_per_band = rpb;
    p.nbands = max(1, (a.height - 1 + rpb - 1) / rpb);

    static unsigned long long attr_mask = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(attr_mask >> dev & 1ULL)) {
        e = cudaFuncSetAttribute(sobel_rows_kernel<CH, FMA_MODE, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_mask |= 1ULL << dev;
    }
    const long long gx = (long long)a.nframes * p.nbands;
    if (gx > 0x7fffffffLL || nstrips > 65535) return cudaErrorInvalidValue;
    count_launch();
    sobel_rows_kernel<CH, FMA_MODE, BIAS><<<dim3((unsigned)gx, (unsigned)nstrips), (p.nw + 1) * 32, smem, st>>>(p);
    return cudaGetLastError();
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
