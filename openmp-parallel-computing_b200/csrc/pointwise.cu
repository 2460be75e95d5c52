// pointwise.cu -- streaming per-pixel kernels: grayscale in place (a1), RGB->luma, plane->RGB
// replicate (a4), legacy (r+g+b)/3 grayscale with min/max (a8), channel sums (a7) and the
// synthetic-input generator.
//
// Data movement (fast path): every warp runs its own ring of shared-memory stages.  One lane
// issues 1-D bulk asynchronous copies (cp.async.bulk, the TMA engine) global->shared, the warp
// transforms the stage in shared memory with conflict-free 128-bit accesses, and the same lane
// issues the bulk copy shared->global.  No LDG/STG on the fast path, so every HBM transaction is
// a full line regardless of the 3-bytes-per-pixel layout; no CTA-wide barrier anywhere.
#include "common.cuh"
#include "kernels.h"

namespace ompimg {

// ---------------------------------------------------------------------------------------------
// per-group (16 pixels) transforms, operating shared -> shared
// ---------------------------------------------------------------------------------------------
template <int FMA_MODE>
struct OpGray3 {   // monolithic/src/parallel_to_grayscale.c:8-16, channels == 3
    static constexpr int IN_GB = 48, OUT_GB = 48;
    static constexpr bool INPLACE = true;
    struct Acc { __device__ void init() {} };
    __device__ static void group(uint32_t in, uint32_t out, Acc &) {
        const uint4 A = lds128(in), B = lds128(in + 16), C = lds128(in + 32);
        const uint32_t w[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
        uint32_t f[16];
        luma16_rgb<FMA_MODE>(w, f, OMPIMG_TWO23);
        uint32_t o[12];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[3 * q] = prmt(f[4 * q], f[4 * q + 1], 0x4000);
            o[3 * q + 1] = prmt(f[4 * q + 1], f[4 * q + 2], 0x4400);
            o[3 * q + 2] = prmt(f[4 * q + 2], f[4 * q + 3], 0x4440);
        }
        sts128(out, make_uint4(o[0], o[1], o[2], o[3]));
        sts128(out + 16, make_uint4(o[4], o[5], o[6], o[7]));
        sts128(out + 32, make_uint4(o[8], o[9], o[10], o[11]));
    }
    __device__ static void pixel(const uint8_t *in, uint8_t *out, Acc &) {
        const uint32_t l = luma1<FMA_MODE>(in[0], in[1], in[2]);
        out[0] = out[1] = out[2] = (uint8_t)l;
    }
};

template <int FMA_MODE>
struct OpGray4 {   // channels == 4: alpha byte untouched (parallel_to_grayscale.c:13-15)
    static constexpr int IN_GB = 64, OUT_GB = 64;
    static constexpr bool INPLACE = true;
    struct Acc { __device__ void init() {} };
    __device__ static void group(uint32_t in, uint32_t out, Acc &) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 v = lds128(in + 16 * i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        uint32_t f[16];
        luma16_rgba<FMA_MODE>(w, f, OMPIMG_TWO23);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = prmt(f[i], w[i], 0x7000);
#pragma unroll
        for (int i = 0; i < 4; ++i) sts128(out + 16 * i, make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
    }
    __device__ static void pixel(const uint8_t *in, uint8_t *out, Acc &) {
        const uint32_t l = luma1<FMA_MODE>(in[0], in[1], in[2]);
        out[0] = out[1] = out[2] = (uint8_t)l;
    }
};

template <int FMA_MODE, int CH>
struct OpLuma {   // RGB(A) -> packed luma plane
    static constexpr int IN_GB = 16 * CH, OUT_GB = 16;
    static constexpr bool INPLACE = false;
    struct Acc { __device__ void init() {} };
    __device__ static void group(uint32_t in, uint32_t out, Acc &) {
        uint32_t f[16];
        if (CH == 3) {
            const uint4 A = lds128(in), B = lds128(in + 16), C = lds128(in + 32);
            const uint32_t w[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
            luma16_rgb<FMA_MODE>(w, f, OMPIMG_TWO23);
        } else {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 v = lds128(in + 16 * i);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
            luma16_rgba<FMA_MODE>(w, f, OMPIMG_TWO23);
        }
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            o[q] = prmt(prmt(f[4 * q], f[4 * q + 1], 0x0040), prmt(f[4 * q + 2], f[4 * q + 3], 0x0040), 0x5410);
        sts128(out, make_uint4(o[0], o[1], o[2], o[3]));
    }
    __device__ static void pixel(const uint8_t *in, uint8_t *out, Acc &) {
        out[0] = (uint8_t)luma1<FMA_MODE>(in[0], in[1], in[2]);
    }
};

struct OpRepl3 {   // monolithic/src/main_with_sobel.c:65-73, channels == 3
    static constexpr int IN_GB = 16, OUT_GB = 48;
    static constexpr bool INPLACE = false;
    struct Acc { __device__ void init() {} };
    __device__ static void group(uint32_t in, uint32_t out, Acc &) {
        const uint4 v = lds128(in);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[12];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[3 * q] = prmt(w[q], w[q], 0x1000);
            o[3 * q + 1] = prmt(w[q], w[q], 0x2211);
            o[3 * q + 2] = prmt(w[q], w[q], 0x3332);
        }
        sts128(out, make_uint4(o[0], o[1], o[2], o[3]));
        sts128(out + 16, make_uint4(o[4], o[5], o[6], o[7]));
        sts128(out + 32, make_uint4(o[8], o[9], o[10], o[11]));
    }
    __device__ static void pixel(const uint8_t *in, uint8_t *out, Acc &) { out[0] = out[1] = out[2] = in[0]; }
};

struct OpAvg3 {   // old/parallel_to_grayscale.c:13-34: v=(r+g+b)/3 to all channels, min/max of v
    static constexpr int IN_GB = 48, OUT_GB = 48;
    static constexpr bool INPLACE = false;
    struct Acc {
        int mn, mx;
        __device__ void init() { mn = 256; mx = -1; }
    };
    __device__ static uint32_t avg(uint32_t r, uint32_t g, uint32_t b, Acc &a) {
        const uint32_t v = ((r + g + b) * 43691u) >> 17;   // exact floor(s/3) for s < 2^16
        a.mn = min(a.mn, (int)v);
        a.mx = max(a.mx, (int)v);
        return v;
    }
    __device__ static void group(uint32_t in, uint32_t out, Acc &acc) {
        const uint4 A = lds128(in), B = lds128(in + 16), C = lds128(in + 32);
        const uint32_t w[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
        uint32_t o[12];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
            const uint32_t v0 = avg(a & 255, (a >> 8) & 255, (a >> 16) & 255, acc);
            const uint32_t v1 = avg(a >> 24, b & 255, (b >> 8) & 255, acc);
            const uint32_t v2 = avg((b >> 16) & 255, b >> 24, c & 255, acc);
            const uint32_t v3 = avg((c >> 8) & 255, (c >> 16) & 255, c >> 24, acc);
            o[3 * q] = v0 * 0x010101u | (v1 << 24);
            o[3 * q + 1] = v1 * 0x0101u | (v2 * 0x01010000u);
            o[3 * q + 2] = v2 | (v3 * 0x01010100u);
        }
        sts128(out, make_uint4(o[0], o[1], o[2], o[3]));
        sts128(out + 16, make_uint4(o[4], o[5], o[6], o[7]));
        sts128(out + 32, make_uint4(o[8], o[9], o[10], o[11]));
    }
    __device__ static void pixel(const uint8_t *in, uint8_t *out, Acc &acc) {
        out[0] = out[1] = out[2] = (uint8_t)avg(in[0], in[1], in[2], acc);
    }
};

template <class Acc>
__device__ __forceinline__ void flush_acc(Acc &, int *) {}
template <>
__device__ __forceinline__ void flush_acc<OpAvg3::Acc>(OpAvg3::Acc &a, int *min_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.mn = min(a.mn, __shfl_xor_sync(0xffffffffu, a.mn, o));
        a.mx = max(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (a.mn < 256) atomicMin(min_max, a.mn);
        if (a.mx >= 0) atomicMax(min_max + 1, a.mx);
    }
}

// ---------------------------------------------------------------------------------------------
// the warp-autonomous bulk pipeline
// ---------------------------------------------------------------------------------------------
constexpr int STREAM_WARPS = 8;
// chunk = 2 groups per lane (3 KB copies for RGB), 3 stages: measured against 1 x 4 / 1 x 6 / 2 x 2 / 2 x 4 / 3 x 2 / 4 x 2 --
// in-place grayscale 976 -> 1029 Gpx/s, RGB->luma 1691 -> 1778, (r+g+b)/3 936 -> 1109 (larger copies matter, deeper rings do not)
#ifndef STREAM_GPL
#define STREAM_GPL 2
#endif
#ifndef STREAM_NSTAGES
#define STREAM_NSTAGES 3
#endif
constexpr int STREAM_STAGES = STREAM_NSTAGES;
constexpr int STREAM_CG = 32 * STREAM_GPL;      // 16-pixel groups per chunk (STREAM_GPL per lane): one bulk copy each way

template <class Op>
constexpr int stream_stage_bytes() { return STREAM_CG * Op::IN_GB + (Op::INPLACE ? 0 : STREAM_CG * Op::OUT_GB); }
template <class Op>
constexpr int stream_smem_bytes() { return STREAM_WARPS * STREAM_STAGES * (stream_stage_bytes<Op>() + 8); }

// A warp visits chunks c_k = gw + k*stride of the (line, chunk-in-line) grid.  Positions advance
// incrementally (no division on the hot path).
struct Walker {
    uint32_t line, cc;
    __device__ __forceinline__ void advance(const StreamParams &p) {
        line += p.step_lines;
        cc += p.step_chunks;
        if (cc >= p.chunks_per_line) { cc -= p.chunks_per_line; ++line; }
    }
};
template <class Op>
__device__ __forceinline__ int chunk_groups(const StreamParams &p, const Walker &w) {
    return (int)min((uint32_t)STREAM_CG, p.groups_per_line - w.cc * (uint32_t)STREAM_CG);
}

template <class Op>
__global__ void __launch_bounds__(STREAM_WARPS * 32) stream_kernel(const StreamParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int S = STREAM_STAGES;
    constexpr int SB = stream_stage_bytes<Op>();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = smem_u32(smem) + warp * (S * SB);
    const uint32_t bars = smem_u32(smem) + STREAM_WARPS * S * SB + warp * (S * 8);

    typename Op::Acc acc;
    acc.init();

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();

    const uint32_t gw = blockIdx.x * STREAM_WARPS + warp;
    const uint32_t stride = gridDim.x * STREAM_WARPS;
    const uint32_t n = p.total_chunks > gw ? (p.total_chunks - gw + stride - 1) / stride : 0;

    Walker cur, nxt;                       // current chunk, next chunk to load
    cur.line = gw / p.chunks_per_line;
    cur.cc = gw - cur.line * p.chunks_per_line;
    nxt = cur;
    uint32_t issued = 0;
    if (lane == 0) {
        for (; issued < (uint32_t)S && issued < n; ++issued) {
            const int ng = chunk_groups<Op>(p, nxt);
            mbar_expect_tx(bars + 8 * issued, ng * Op::IN_GB);
            bulk_g2s(base + issued * SB, p.in + (size_t)nxt.line * p.in_line_pitch + (size_t)nxt.cc * (STREAM_CG * Op::IN_GB),
                     ng * Op::IN_GB, bars + 8 * issued);
            nxt.advance(p);
        }
    }
    int s = 0;
    uint32_t ph = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const int ng = chunk_groups<Op>(p, cur);
        const uint32_t in_s = base + s * SB;
        const uint32_t out_s = Op::INPLACE ? in_s : in_s + STREAM_CG * Op::IN_GB;
        mbar_wait(bars + 8 * s, ph);
#pragma unroll
        for (int gl = 0; gl < STREAM_GPL; ++gl) {
            const int gi = lane + 32 * gl;
            if (gi < ng) Op::group(in_s + gi * Op::IN_GB, out_s + gi * Op::OUT_GB, acc);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(p.out + (size_t)cur.line * p.out_line_pitch + (size_t)cur.cc * (STREAM_CG * Op::OUT_GB), out_s, ng * Op::OUT_GB);
            bulk_commit();
            // refill the stage whose store was issued one step ago (its smem reads must be done)
            if (k >= 1 && issued < n) {
                bulk_wait_read<1>();
                const int sn = (s == 0) ? S - 1 : s - 1;
                const int ngn = chunk_groups<Op>(p, nxt);
                mbar_expect_tx(bars + 8 * sn, ngn * Op::IN_GB);
                bulk_g2s(base + sn * SB, p.in + (size_t)nxt.line * p.in_line_pitch + (size_t)nxt.cc * (STREAM_CG * Op::IN_GB),
                         ngn * Op::IN_GB, bars + 8 * sn);
                nxt.advance(p);
                ++issued;
            }
        }
        cur.advance(p);
        if (++s == S) { s = 0; ph ^= 1; }
    }
    if (lane == 0) bulk_wait_all<0>();

    // pixels beyond the last full 16-pixel group of every line: byte path
    if (p.tail_px > 0) {
        const long long nt = (long long)p.n_lines * p.tail_px;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += (long long)gridDim.x * blockDim.x) {
            const long long line = i / p.tail_px;
            const int t = (int)(i - line * p.tail_px);
            const long long px = (long long)p.groups_per_line * 16 + t;
            Op::pixel(p.in + line * p.in_line_pitch + px * (Op::IN_GB / 16),
                      p.out + line * p.out_line_pitch + px * (Op::OUT_GB / 16), acc);
        }
    }
    flush_acc(acc, p.min_max);
}

// generic fallback: one thread per pixel, byte accesses, any alignment / pitch
template <class Op>
__global__ void __launch_bounds__(256) generic_kernel(const StreamParams p, long long px_per_line) {
    typename Op::Acc acc;
    acc.init();
    const long long total = (long long)p.n_lines * px_per_line;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long line = i / px_per_line;
        const long long x = i - line * px_per_line;
        Op::pixel(p.in + line * p.in_line_pitch + x * (Op::IN_GB / 16),
                  p.out + line * p.out_line_pitch + x * (Op::OUT_GB / 16), acc);
    }
    flush_acc(acc, p.min_max);
}

template <class Op>
static cudaError_t set_smem_attr_once(int smem) {
    static unsigned long long done_mask = 0;   // per Op instantiation, one bit per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(done_mask >> dev & 1ULL)) {
        e = cudaFuncSetAttribute(stream_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        done_mask |= 1ULL << dev;
    }
    return cudaSuccess;
}

// One launch over `n_lines` lines of `px_per_line` pixels; line i starts at in + i*in_line_pitch.
template <class Op>
static cudaError_t launch_lines(const uint8_t *in, uint8_t *out, size_t in_line_pitch, size_t out_line_pitch, long long n_lines,
                                long long px_per_line, int *min_max, int sm_count, cudaStream_t st) {
    if (n_lines <= 0 || px_per_line <= 0) return cudaSuccess;
    StreamParams p{};
    p.in = in; p.out = out; p.in_line_pitch = in_line_pitch; p.out_line_pitch = out_line_pitch; p.min_max = min_max;
    p.n_lines = n_lines;
    const bool fast = (((uintptr_t)in | (uintptr_t)out) & 15) == 0 &&
                      (n_lines == 1 || ((in_line_pitch | out_line_pitch) & 15) == 0);
    const long long groups = px_per_line / 16, cpl = (groups + STREAM_CG - 1) / STREAM_CG;
    if (fast && n_lines * (cpl > 0 ? cpl : 1) < (1LL << 31) && groups < (1LL << 31)) {
        p.groups_per_line = (uint32_t)groups;
        p.tail_px = (int)(px_per_line % 16);
        p.chunks_per_line = (uint32_t)(cpl > 0 ? cpl : 1);
        p.total_chunks = (uint32_t)(cpl > 0 ? n_lines * cpl : 0);
        const int smem = stream_smem_bytes<Op>();
        cudaError_t e = set_smem_attr_once<Op>(smem);
        if (e != cudaSuccess) return e;
        const int ctas_per_sm = max(1, min(4, (220 * 1024) / (smem + 1024)));
        long long want = ((long long)p.total_chunks + STREAM_WARPS - 1) / STREAM_WARPS;
        const long long tail_ctas = (n_lines * p.tail_px + 255) / 256;
        if (want < tail_ctas) want = tail_ctas;
        if (want < 1) want = 1;
        const int grid = (int)min(want, (long long)sm_count * ctas_per_sm);
        const uint32_t stride = (uint32_t)grid * STREAM_WARPS;
        p.step_lines = stride / p.chunks_per_line;
        p.step_chunks = stride % p.chunks_per_line;
        count_launch();
        stream_kernel<Op><<<grid, STREAM_WARPS * 32, smem, st>>>(p);
    } else {
        const long long total = n_lines * px_per_line;
        const int grid = (int)max(1LL, min((total + 255) / 256, (long long)sm_count * 16));
        count_launch();
        generic_kernel<Op><<<grid, 256, 0, st>>>(p, px_per_line);
    }
    return cudaGetLastError();
}

// frames x rows with arbitrary pitches/strides -> as few launches of launch_lines as the layout allows
template <class Op>
static cudaError_t launch_image(const uint8_t *in, uint8_t *out, size_t in_pitch, size_t out_pitch, size_t in_fstride, size_t out_fstride,
                                int w, int h, int nframes, int *min_max, int sm_count, cudaStream_t st) {
    const size_t in_row = (size_t)w * (Op::IN_GB / 16), out_row = (size_t)w * (Op::OUT_GB / 16);
    if (in_pitch == in_row && out_pitch == out_row) {
        // dense rows: a frame is one line
        if (nframes == 1 || (in_fstride == in_row * h && out_fstride == out_row * h))
            return launch_lines<Op>(in, out, in_row * h * nframes, out_row * h * nframes, 1, (long long)w * h * nframes, min_max, sm_count, st);
        return launch_lines<Op>(in, out, in_fstride, out_fstride, nframes, (long long)w * h, min_max, sm_count, st);
    }
    if (nframes == 1 || (in_fstride == in_pitch * h && out_fstride == out_pitch * h))
        return launch_lines<Op>(in, out, in_pitch, out_pitch, (long long)h * nframes, w, min_max, sm_count, st);
    for (int f = 0; f < nframes; ++f) {
        cudaError_t e = launch_lines<Op>(in + f * in_fstride, out + f * out_fstride, in_pitch, out_pitch, h, w, min_max, sm_count, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// host-callable launchers (see kernels.h)
// ---------------------------------------------------------------------------------------------
cudaError_t launch_gray_inplace(uint8_t *img, size_t pitch, int w, int h, int ch, int nframes, size_t fstride,
                                int fma_mode, int sm_count, cudaStream_t st) {
    if (ch == 3)
        return fma_mode ? launch_image<OpGray3<1>>(img, img, pitch, pitch, fstride, fstride, w, h, nframes, nullptr, sm_count, st)
                        : launch_image<OpGray3<0>>(img, img, pitch, pitch, fstride, fstride, w, h, nframes, nullptr, sm_count, st);
    return fma_mode ? launch_image<OpGray4<1>>(img, img, pitch, pitch, fstride, fstride, w, h, nframes, nullptr, sm_count, st)
                    : launch_image<OpGray4<0>>(img, img, pitch, pitch, fstride, fstride, w, h, nframes, nullptr, sm_count, st);
}

cudaError_t launch_rgb_to_luma(const uint8_t *rgb, size_t rgb_pitch, int ch, uint8_t *luma, size_t luma_pitch, int w, int h,
                               int nframes, size_t rgb_stride, size_t luma_stride, int fma_mode, int sm_count, cudaStream_t st) {
    if (ch == 3)
        return fma_mode ? launch_image<OpLuma<1, 3>>(rgb, luma, rgb_pitch, luma_pitch, rgb_stride, luma_stride, w, h, nframes, nullptr, sm_count, st)
                        : launch_image<OpLuma<0, 3>>(rgb, luma, rgb_pitch, luma_pitch, rgb_stride, luma_stride, w, h, nframes, nullptr, sm_count, st);
    return fma_mode ? launch_image<OpLuma<1, 4>>(rgb, luma, rgb_pitch, luma_pitch, rgb_stride, luma_stride, w, h, nframes, nullptr, sm_count, st)
                    : launch_image<OpLuma<0, 4>>(rgb, luma, rgb_pitch, luma_pitch, rgb_stride, luma_stride, w, h, nframes, nullptr, sm_count, st);
}

// channels == 4 keeps alpha, so it cannot overwrite whole groups: per-pixel kernel
__global__ void __launch_bounds__(256) replicate4_kernel(const uint8_t *plane, size_t ppitch, uint8_t *img, size_t ipitch, int w, int h,
                                                        int nframes, size_t pstride, size_t istride) {
    const long long total = (long long)nframes * h * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long rl = i / w;
        const long long x = i - rl * w, f = rl / h, y = rl - f * h;
        const uint8_t e = plane[f * pstride + y * ppitch + x];
        uint8_t *o = img + f * istride + y * ipitch + x * 4;
        o[0] = e; o[1] = e; o[2] = e;
    }
}

cudaError_t launch_replicate(const uint8_t *plane, size_t ppitch, uint8_t *img, size_t ipitch, int ch, int w, int h, int nframes,
                             size_t pstride, size_t istride, int sm_count, cudaStream_t st) {
    if (ch == 4) {
        const long long total = (long long)nframes * h * w;
        const int grid = (int)max(1LL, min((total + 255) / 256, (long long)sm_count * 16));
        count_launch();
        replicate4_kernel<<<grid, 256, 0, st>>>(plane, ppitch, img, ipitch, w, h, nframes, pstride, istride);
        return cudaGetLastError();
    }
    return launch_image<OpRepl3>(plane, img, ppitch, ipitch, pstride, istride, w, h, nframes, nullptr, sm_count, st);
}

cudaError_t launch_gray_avg3(const uint8_t *rgb, size_t in_pitch, uint8_t *out, size_t out_pitch, int rows, int cols, int *min_max_dev,
                             int sm_count, cudaStream_t st) {
    return launch_image<OpAvg3>(rgb, out, in_pitch, out_pitch, 0, 0, cols, rows, 1, min_max_dev, sm_count, st);
}

// ---------------------------------------------------------------------------------------------
// a7: channel sums (old/parallel_avg_pixel.c:17-30).  Pure read stream, 3 B/px.
//
// A "line" is a run of whole pixels that starts 16-byte aligned (the whole image when rows are
// dense, else one row).  A warp reads tiles of 96 consecutive 16-byte pieces with three fully
// coalesced 128-bit loads (lane t takes pieces t, t+32, t+64 of the tile).  Piece j of a line
// starts at byte 16j, i.e. at channel phase j mod 3 -- and (t + 32k) mod 3 does not depend on the
// tile, so a lane simply adds up the bytes of each piece by byte position mod 3 (12 dp4a against
// constant masks, no shuffling of bytes) into 3x3 accumulators and rotates them to channels once,
// at the end.  Two tiles are in flight per warp iteration.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// sums of the bytes at positions = 0, 1, 2 (mod 3) of a 16-byte piece
__device__ __forceinline__ void piece_sums(const uint4 v, uint32_t (&a)[3]) {
    // byte position of word w, byte b is 4w+b: class (4w+b) mod 3 = (w+b) mod 3
    a[0] = dp4a_uu(v.x, 0x01000001u, a[0]); a[1] = dp4a_uu(v.x, 0x00000100u, a[1]); a[2] = dp4a_uu(v.x, 0x00010000u, a[2]);
    a[1] = dp4a_uu(v.y, 0x01000001u, a[1]); a[2] = dp4a_uu(v.y, 0x00000100u, a[2]); a[0] = dp4a_uu(v.y, 0x00010000u, a[0]);
    a[2] = dp4a_uu(v.z, 0x01000001u, a[2]); a[0] = dp4a_uu(v.z, 0x00000100u, a[0]); a[1] = dp4a_uu(v.z, 0x00010000u, a[1]);
    a[0] = dp4a_uu(v.w, 0x01000001u, a[0]); a[1] = dp4a_uu(v.w, 0x00000100u, a[1]); a[2] = dp4a_uu(v.w, 0x00010000u, a[2]);
}

__global__ void __launch_bounds__(256) channel_sums_kernel(const uint8_t *rgb, size_t line_pitch, long long nlines, long long line_bytes,
                                                           unsigned long long *sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pieces = line_bytes >> 4;                    // whole 16-byte pieces per line
    const long long tiles = (pieces + 95) / 96;
    uint32_t acc[3][3];                                          // [k: which of the lane's three pieces][byte position mod 3]
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k][0] = acc[k][1] = acc[k][2] = 0u;
    unsigned long long s[3] = {0, 0, 0};                         // channel sums of this thread
    const uint4 Z = make_uint4(0u, 0u, 0u, 0u);
    for (long long line = blockIdx.y; line < nlines; line += gridDim.y) {
        const uint4 *base = reinterpret_cast<const uint4 *>(rgb + line * line_pitch);
        const long long tstep = (long long)gridDim.x * 8;
        for (long long t = (long long)blockIdx.x * 8 + warp; t < tiles; t += 2 * tstep) {
            const long long j0 = t * 96 + lane, j1 = j0 + tstep * 96;
            uint4 v[6];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v[k] = j0 + 32 * k < pieces ? __ldg(base + j0 + 32 * k) : Z;
                v[3 + k] = j1 + 32 * k < pieces ? __ldg(base + j1 + 32 * k) : Z;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { piece_sums(v[k], acc[k]); piece_sums(v[3 + k], acc[k]); }
        }
        // the < 16 bytes after the last whole piece of the line
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const uint8_t *q = rgb + line * line_pitch;
            for (long long i = pieces << 4; i < line_bytes; ++i) {
                const int c = (int)(i % 3);
                s[0] += c == 0 ? q[i] : 0; s[1] += c == 1 ? q[i] : 0; s[2] += c == 2 ? q[i] : 0;
            }
        }
    }
    // piece k of lane t sits at channel phase (t + 32k) mod 3: byte position class c holds channel (c + phase) mod 3
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int phase = (lane + 2 * k) % 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int ch = (c + phase) % 3;
            s[0] += ch == 0 ? acc[k][c] : 0u;
            s[1] += ch == 1 ? acc[k][c] : 0u;
            s[2] += ch == 2 ? acc[k][c] : 0u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
        s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
        s[2] += __shfl_xor_sync(0xffffffffu, s[2], o);
    }
    // one atomic triple per CTA: thousands of same-address atomics from every warp cost more than the read itself
    __shared__ unsigned long long part[8][3];
    if (lane == 0) { part[warp][0] = s[0]; part[warp][1] = s[1]; part[warp][2] = s[2]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += part[w][threadIdx.x];
        atomicAdd(sums + threadIdx.x, tot);
    }
}

// any alignment / pitch: one thread per pixel
__global__ void __launch_bounds__(256) channel_sums_generic_kernel(const uint8_t *rgb, size_t pitch, int rows, int cols, unsigned long long *sums) {
    unsigned long long s0 = 0, s1 = 0, s2 = 0;
    const long long total = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, x = i - r * cols;
        const uint8_t *px = rgb + r * pitch + x * 3;
        s0 += px[0]; s1 += px[1]; s2 += px[2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(sums, s0); atomicAdd(sums + 1, s1); atomicAdd(sums + 2, s2); }
}

cudaError_t launch_channel_sums(const uint8_t *rgb, size_t pitch, int rows, int cols, unsigned long long *sums_dev, int sm_count, cudaStream_t st) {
    const size_t row_bytes = (size_t)cols * 3;
    const bool dense = pitch == row_bytes || rows == 1;
    count_launch();
    if (((uintptr_t)rgb & 15) == 0 && (dense || (pitch & 15) == 0)) {
        const long long nlines = dense ? 1 : rows;
        const long long line_bytes = dense ? (long long)row_bytes * rows : (long long)row_bytes;
        // accumulators are 32-bit: 6 * 255 per piece, a lane sees line_bytes / (16 * 32 * 8 * grid.x) pieces per k
        const long long tiles = (line_bytes / 16 + 95) / 96;
        const long long gx = max(1LL, min((tiles + 15) / 16, dense ? (long long)sm_count * 8 : 64LL));
        const long long gy = max(1LL, min(nlines, ((long long)sm_count * 8 + gx - 1) / gx));
        // worst case per 32-bit accumulator: pieces per lane and k over the whole launch
        const double per_lane = (double)line_bytes / 16.0 / 96.0 / (double)(gx * 8) * (double)((nlines + gy - 1) / gy) + 2.0;
        if (per_lane * 6.0 * 255.0 < 4.0e9) {
            channel_sums_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(rgb, pitch, nlines, line_bytes, sums_dev);
            return cudaGetLastError();
        }
    }
    const long long total = (long long)rows * cols;
    const int grid = (int)max(1LL, min((total + 255) / 256, (long long)sm_count * 16));
    channel_sums_generic_kernel<<<grid, 256, 0, st>>>(rgb, pitch, rows, cols, sums_dev);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// synthetic input: byte i = mix32(seed + (i>>2)*0x9E3779B9) >> 8*(i&3)   (same as oracle_synth_fill)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__global__ void __launch_bounds__(256) synth_kernel(uint8_t *dst, unsigned long long nbytes, uint32_t seed, unsigned long long off) {
    // head bytes up to 16-byte alignment of (dst), body as uint4, tail bytes
    const unsigned long long head = min(nbytes, (unsigned long long)((16 - ((uintptr_t)dst & 15)) & 15));
    const unsigned long long nvec = (nbytes - head) / 16;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
    if (((off + head) & 3) == 0) {
        uint4 *v = reinterpret_cast<uint4 *>(dst + head);
        const unsigned long long w0 = (off + head) >> 2;
        for (unsigned long long i = tid; i < nvec; i += nth) {
            const uint32_t j = (uint32_t)(w0 + 4 * i);
            v[i] = make_uint4(mix32(seed + j * 0x9E3779B9u), mix32(seed + (j + 1) * 0x9E3779B9u),
                              mix32(seed + (j + 2) * 0x9E3779B9u), mix32(seed + (j + 3) * 0x9E3779B9u));
        }
    } else {
        for (unsigned long long i = tid; i < nvec * 16; i += nth) {
            const unsigned long long gi = off + head + i;
            dst[head + i] = (uint8_t)(mix32(seed + (uint32_t)(gi >> 2) * 0x9E3779B9u) >> (8 * (gi & 3)));
        }
    }
    for (unsigned long long i = tid; i < head; i += nth) {
        const unsigned long long gi = off + i;
        dst[i] = (uint8_t)(mix32(seed + (uint32_t)(gi >> 2) * 0x9E3779B9u) >> (8 * (gi & 3)));
    }
    const unsigned long long t0 = head + nvec * 16;
    for (unsigned long long i = t0 + tid; i < nbytes; i += nth) {
        const unsigned long long gi = off + i;
        dst[i] = (uint8_t)(mix32(seed + (uint32_t)(gi >> 2) * 0x9E3779B9u) >> (8 * (gi & 3)));
    }
}
cudaError_t launch_synth(uint8_t *dst, unsigned long long nbytes, unsigned int seed, unsigned long long off, int sm_count, cudaStream_t st) {
    const unsigned long long nvec = nbytes / 16 + 1;
    const int grid = (int)max(1ULL, min((nvec + 255) / 256, (unsigned long long)sm_count * 16));
    count_launch();
    synth_kernel<<<grid, 256, 0, st>>>(dst, nbytes, seed, off);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// self test: the device luma function over all 2^24 triples
// ---------------------------------------------------------------------------------------------
template <int FMA_MODE>
__global__ void luma_lut_kernel(uint8_t *lut) {
    // thread -> 16 consecutive triples (r,g fixed, b0..b0+15), run through the SAME packed path as the image kernels
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;   // 2^20 threads
    const uint32_t r = i >> 12, g = (i >> 4) & 255, b0 = (i & 15) * 16;
    uint8_t px[48];
#pragma unroll
    for (int k = 0; k < 16; ++k) { px[3 * k] = (uint8_t)r; px[3 * k + 1] = (uint8_t)g; px[3 * k + 2] = (uint8_t)(b0 + k); }
    uint32_t w[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) w[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
    uint32_t f[16];
    luma16_rgb<FMA_MODE>(w, f, OMPIMG_TWO23);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t l = f[k] & 255u;
        lut[(r << 16) | (g << 8) | (b0 + k)] = (uint8_t)l;
    }
}
cudaError_t launch_luma_lut(uint8_t *lut_dev, int fma_mode, cudaStream_t st) {
    if (fma_mode) { count_launch(); luma_lut_kernel<1><<<(1 << 20) / 256, 256, 0, st>>>(lut_dev); }
    else { count_launch(); luma_lut_kernel<0><<<(1 << 20) / 256, 256, 0, st>>>(lut_dev); }
    return cudaGetLastError();
}

}  // namespace ompimg
