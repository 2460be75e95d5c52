// common.cuh -- device-side building blocks shared by every kernel of libompimg_b200.
//
// Everything here is sm_100a-only: bulk asynchronous copies (cp.async.bulk -> UBLKCP) with
// mbarrier completion, packed f32x2 arithmetic (FFMA2/FADD2), packed f16x2 arithmetic
// (HFMA2/HADD2), mixed-precision FHFMA, MUFU.SQRT and the saturating F2IP pack.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ompimg {

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// byte permute / packed helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ u64 pack64(uint32_t lo, uint32_t hi) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ void unpack64(u64 v, uint32_t &lo, uint32_t &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ u64 splat2(float x) { return pack64(__float_as_uint(x), __float_as_uint(x)); }

// pin a loop-invariant value in a register (ptxas cannot rematerialise what it cannot see through)
__device__ __forceinline__ uint32_t keep(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ u64 keep64(u64 x) { asm volatile("" : "+l"(x)); return x; }

// packed fp32 pairs (one issue slot for two pixels)
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fadd2_rz(u64 a, u64 b) {
    u64 d;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// packed fp16 pairs
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// m = gx.h^2 + gy.h^2 + c in fp32 from one half of two f16x2 registers (FHFMA x2, exact for
// integer-valued halves up to 2048)
template <int HI>
__device__ __forceinline__ float sqsum_f16(uint32_t gx, uint32_t gy, float c) {
    float m;
    if (HI == 0) {
        asm("{\n .reg .f16 xl, xh, yl, yh;\n mov.b32 {xl, xh}, %1;\n mov.b32 {yl, yh}, %2;\n"
            " fma.rn.f32.f16 %0, xl, xl, %3;\n fma.rn.f32.f16 %0, yl, yl, %0;\n}"
            : "=f"(m) : "r"(gx), "r"(gy), "f"(c));
    } else {
        asm("{\n .reg .f16 xl, xh, yl, yh;\n mov.b32 {xl, xh}, %1;\n mov.b32 {yl, yh}, %2;\n"
            " fma.rn.f32.f16 %0, xh, xh, %3;\n fma.rn.f32.f16 %0, yh, yh, %0;\n}"
            : "=f"(m) : "r"(gx), "r"(gy), "f"(c));
    }
    return m;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// d = (c << 16) | (sat_u8(trunc(hi)) << 8) | sat_u8(trunc(lo))   -> one F2IP.U8.F32.TRUNC
__device__ __forceinline__ uint32_t pack_sat_u8(float hi, float lo, uint32_t c) {
    uint32_t d;
    asm("{\n .reg .s32 a, b;\n cvt.rzi.s32.f32 a, %1;\n cvt.rzi.s32.f32 b, %2;\n"
        " cvt.pack.sat.u8.s32.b32 %0, a, b, %3;\n}"
        : "=r"(d) : "f"(hi), "f"(lo), "r"(c));
    return d;
}

// ------------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA engine, 1-D form)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    // the suspend-time hint lets the hardware park the warp until the phase completes instead of
    // returning early and burning issue slots in the retry loop
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// for a thread that is usually far ahead of the event it waits for (a producer waiting for a free stage):
// sleep between probes instead of competing for issue slots with the warps doing the work
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(256);
}
// global -> shared, completion counted in bytes on `bar`.  src, dst 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk group
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (before a bulk store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// luma arithmetic (monolithic/src/parallel_to_grayscale.c:12)
//
// A channel byte c is unpacked with one PRMT into the fp32 bit pattern 0x4B0000cc = 2^23 + c.
// strict:  RN(k*c) = fma(k, 2^23+c, -k*2^23): the product-sum inside the FMA is exactly k*c, so
//          one rounding gives the same bits as the reference's `k*(float)c` (mulss), then two
//          rounded adds in the reference's order ((r+g)+b) and truncation.
// fma:     gcc -march=native contracts the expression to fma(kb,b, fma(kr,r, RN(kg*g)))
//          (SURVEY appendix A.1); r and b are made exact floats with one add each.
// truncation: add.rz(v, MAGIC) with MAGIC in [2^23,2^24) leaves floor(v) in the low mantissa byte.
// ------------------------------------------------------------------------------------------------
#define OMPIMG_KR 0.299f
#define OMPIMG_KG 0.587f
#define OMPIMG_KB 0.114f
#define OMPIMG_TWO23 8388608.0f

// two pixels at once; inputs are bit patterns 0x4B0000cc packed in pairs; returns the pair of
// fp32 bit patterns (MAGIC + luma), luma in the low byte of each half.
template <int FMA_MODE>
__device__ __forceinline__ u64 luma2(u64 xr, u64 xg, u64 xb, float magic) {
    const u64 KR = splat2(OMPIMG_KR), KG = splat2(OMPIMG_KG), KB = splat2(OMPIMG_KB);
    const u64 M = splat2(magic);
    if (FMA_MODE == 0) {
        u64 pr = ffma2(KR, xr, splat2(-OMPIMG_KR * OMPIMG_TWO23));
        u64 pg = ffma2(KG, xg, splat2(-OMPIMG_KG * OMPIMG_TWO23));
        u64 pb = ffma2(KB, xb, splat2(-OMPIMG_KB * OMPIMG_TWO23));
        return fadd2_rz(fadd2(fadd2(pr, pg), pb), M);
    } else {
        const u64 NEG = splat2(-OMPIMG_TWO23);
        u64 pg = ffma2(KG, xg, splat2(-OMPIMG_KG * OMPIMG_TWO23));
        u64 t = ffma2(KR, fadd2(xr, NEG), pg);
        u64 v = ffma2(KB, fadd2(xb, NEG), t);
        return fadd2_rz(v, M);
    }
}

// scalar form for edge pixels and generic kernels (same bits as luma2)
template <int FMA_MODE>
__device__ __forceinline__ uint32_t luma1(uint32_t r, uint32_t g, uint32_t b) {
    float fr = (float)r, fg = (float)g, fb = (float)b;
    float v;
    if (FMA_MODE == 0) v = __fadd_rn(__fadd_rn(__fmul_rn(OMPIMG_KR, fr), __fmul_rn(OMPIMG_KG, fg)), __fmul_rn(OMPIMG_KB, fb));
    else v = __fmaf_rn(OMPIMG_KB, fb, __fmaf_rn(OMPIMG_KR, fr, __fmul_rn(OMPIMG_KG, fg)));
    return (uint32_t)(int)v;   // v in [0,255]: truncation toward zero
}

// one pixel whose r,g,b are bytes 0..2 of `px` -> fp32 bit pattern (MAGIC + luma); same arithmetic as luma2
template <int FMA_MODE>
__device__ __forceinline__ uint32_t luma_px(uint32_t px, float magic) {
    const uint32_t Z = 0x4B000000u;
    const float xr = __uint_as_float(prmt(px, Z, 0x7440)), xg = __uint_as_float(prmt(px, Z, 0x7441)), xb = __uint_as_float(prmt(px, Z, 0x7442));
    float v;
    if (FMA_MODE == 0) {
        const float pr = __fmaf_rn(OMPIMG_KR, xr, -OMPIMG_KR * OMPIMG_TWO23);
        const float pg = __fmaf_rn(OMPIMG_KG, xg, -OMPIMG_KG * OMPIMG_TWO23);
        const float pb = __fmaf_rn(OMPIMG_KB, xb, -OMPIMG_KB * OMPIMG_TWO23);
        v = __fadd_rn(__fadd_rn(pr, pg), pb);
    } else {
        const float pg = __fmaf_rn(OMPIMG_KG, xg, -OMPIMG_KG * OMPIMG_TWO23);
        const float t = __fmaf_rn(OMPIMG_KR, __fadd_rn(xr, -OMPIMG_TWO23), pg);
        v = __fmaf_rn(OMPIMG_KB, __fadd_rn(xb, -OMPIMG_TWO23), t);
    }
    return __float_as_uint(__fadd_rz(v, magic));
}

// 16 pixels of interleaved RGB (12 words) -> 16 fp32 bit patterns (MAGIC + luma)
template <int FMA_MODE>
__device__ __forceinline__ void luma16_rgb(const uint32_t (&w)[12], uint32_t (&f)[16], float magic) {
    const uint32_t Z = 0x4B000000u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // word a = (r0 g0 b0 r1)  b = (g1 b1 r2 g2)  c = (b2 r3 g3 b3)
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        uint32_t r0 = prmt(a, Z, 0x7440), g0 = prmt(a, Z, 0x7441), b0 = prmt(a, Z, 0x7442), r1 = prmt(a, Z, 0x7443);
        uint32_t g1 = prmt(b, Z, 0x7440), b1 = prmt(b, Z, 0x7441), r2 = prmt(b, Z, 0x7442), g2 = prmt(b, Z, 0x7443);
        uint32_t b2 = prmt(c, Z, 0x7440), r3 = prmt(c, Z, 0x7441), g3 = prmt(c, Z, 0x7442), b3 = prmt(c, Z, 0x7443);
        u64 v01 = luma2<FMA_MODE>(pack64(r0, r1), pack64(g0, g1), pack64(b0, b1), magic);
        u64 v23 = luma2<FMA_MODE>(pack64(r2, r3), pack64(g2, g3), pack64(b2, b3), magic);
        unpack64(v01, f[4 * q], f[4 * q + 1]);
        unpack64(v23, f[4 * q + 2], f[4 * q + 3]);
    }
}
// 16 pixels of interleaved RGBA (16 words)
template <int FMA_MODE>
__device__ __forceinline__ void luma16_rgba(const uint32_t (&w)[16], uint32_t (&f)[16], float magic) {
    const uint32_t Z = 0x4B000000u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint32_t a = w[2 * q], b = w[2 * q + 1];
        uint32_t r0 = prmt(a, Z, 0x7440), g0 = prmt(a, Z, 0x7441), b0 = prmt(a, Z, 0x7442);
        uint32_t r1 = prmt(b, Z, 0x7440), g1 = prmt(b, Z, 0x7441), b1 = prmt(b, Z, 0x7442);
        u64 v = luma2<FMA_MODE>(pack64(r0, r1), pack64(g0, g1), pack64(b0, b1), magic);
        unpack64(v, f[2 * q], f[2 * q + 1]);
    }
}

// ------------------------------------------------------------------------------------------------
// Sobel arithmetic (monolithic/src/sobel.c:14-21) on f16x2 pairs.
//
// A pixel value v is carried as the half 0x6400|v = 1024+v (one PRMT from a byte).  With
// P_k=(v[2k],v[2k+1]) and Q_k=(v[2k-1],v[2k]) for one image row:
//   Dh_k  = Q_{k+1}-Q_k              = v[x+1]-v[x-1]                       (bias cancels)
//   Sh'_k = (2*P_k-3072)+Q_k+Q_{k+1} = v[x-1]+2v[x]+v[x+1]+1024            (every step exact, |.|<=2048)
// and over three rows a,b,c:  gx = Dh_a+2Dh_b+Dh_c = (Dh_a+Dh_b)+(Dh_b+Dh_c),  gy = Sh'_a-Sh'_c,
// which are exactly the reference's gx, gy (|g| <= 1020).  m = gx^2+gy^2+0.5 is exact in fp32
// (FHFMA), sqrt.approx(m) is within 2^-22 relative, so its truncation equals isqrt(gx^2+gy^2)
// whenever that is < 256 (distance from an integer >= 0.25/255), and F2IP saturates the rest to
// 255: the same value as (int)sqrtf(float) clamped to 255.
// ------------------------------------------------------------------------------------------------
struct SobelRowState {   // per 2 pixels
    uint32_t sab;   // Dh(row a) + Dh(row b)
    uint32_t dhb;   // Dh(row b)
    uint32_t sha;   // Sh'(row a)
    uint32_t shb;   // Sh'(row b)
};

// horizontal pass of a new row c for one pixel pair, then the vertical combination with the
// two previous rows.  Returns gx, gy of the middle row b and rotates the state.
__device__ __forceinline__ void sobel_pair(uint32_t P, uint32_t Qk, uint32_t Qk1, SobelRowState &st,
                                           uint32_t &gx, uint32_t &gy) {
    const uint32_t dh = hsub2(Qk1, Qk);
    const uint32_t t1 = hfma2(P, 0x40004000u /*2,2*/, 0xEA00EA00u /*-3072,-3072*/);
    const uint32_t sh = hadd2(hadd2(t1, Qk), Qk1);
    const uint32_t sbc = hadd2(st.dhb, dh);
    gx = hadd2(st.sab, sbc);
    gy = hsub2(st.sha, sh);
    st.sab = sbc;
    st.dhb = dh;
    st.sha = st.shb;
    st.shb = sh;
}

// magnitudes of 8 pixel pairs -> 16 packed bytes
__device__ __forceinline__ uint4 sobel_mag16(const uint32_t (&gx)[8], const uint32_t (&gy)[8], float half) {
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float s0 = sqrt_approx(sqsum_f16<0>(gx[2 * j], gy[2 * j], half));
        float s1 = sqrt_approx(sqsum_f16<1>(gx[2 * j], gy[2 * j], half));
        float s2 = sqrt_approx(sqsum_f16<0>(gx[2 * j + 1], gy[2 * j + 1], half));
        float s3 = sqrt_approx(sqsum_f16<1>(gx[2 * j + 1], gy[2 * j + 1], half));
        o[j] = pack_sat_u8(s1, s0, pack_sat_u8(s3, s2, 0u));
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// store the first `n` (0..16) bytes of v at dst+skip.. : bytes [skip, skip+n) of the 16-byte group
// at 16-byte aligned dst.  Used for the groups that touch the image border, which must keep
// the destination's border bytes (sobel.c:11-12 never writes them).
__device__ __forceinline__ void store_partial16(uint8_t *dst, uint4 v, int first, int last /*exclusive*/) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (i >= first && i < last) dst[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
}

// the two border cases of a 16-byte output group, with few wide stores instead of 16 byte stores
//   mode 2: the group holds x == 0          -> bytes 1..15
//   mode 3: the group holds x == width-1    -> bytes 0..plast-1 (plast in 0..15)
//   mode 4: both (image narrower than 17)   -> byte loop
// (forced inline: as a called function its ABI pinned the row state of the Sobel kernels to fixed registers and cost
// ~20 register moves per row on the hot path -- +9-11 % when inlined, although the code itself runs for two lanes per row)
static __device__ __forceinline__ void store_edge16(uint8_t *q, uint4 o, int mode, int plast) {
    if (mode == 2) {
        q[1] = (uint8_t)(o.x >> 8);
        *reinterpret_cast<uint16_t *>(q + 2) = (uint16_t)(o.x >> 16);
        *reinterpret_cast<uint32_t *>(q + 4) = o.y;
        *reinterpret_cast<uint2 *>(q + 8) = make_uint2(o.z, o.w);
    } else if (mode == 3) {
        const int n = plast;
        if (n & 8) *reinterpret_cast<uint2 *>(q) = make_uint2(o.x, o.y);
        uint32_t w = (n & 8) ? o.z : o.x;
        const uint32_t wn = (n & 8) ? o.w : o.y;
        int pos = n & 8;
        if (n & 4) { *reinterpret_cast<uint32_t *>(q + pos) = w; pos += 4; w = wn; }
        if (n & 2) { *reinterpret_cast<uint16_t *>(q + pos) = (uint16_t)w; pos += 2; w >>= 16; }
        if (n & 1) q[pos] = (uint8_t)w;
    } else {
        store_partial16(q, o, 1, plast);
    }
}

}  // namespace ompimg
