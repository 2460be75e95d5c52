/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product library.
 *
 * CPU restatement of the per-pixel hot path of
 * PedemonteGiacomo/OpenMp-Parallel-Computing, written from the behaviour of the
 * reference sources (cited per function, paths relative to the reference root).
 * It exists so that the CUDA path can be checked bit-for-bit on machines where
 * the reference sources are not present (the GPU box).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product never does.
 *
 * Pinning: `make -C oracle check` (= tests/test_oracle.py) compare every
 * function here against the UNMODIFIED reference sources compiled into
 * oracle/_ref/ (see Makefile), and the grayscale path against the reference's one
 * golden vector images/test_gray.png.
 *
 * Differences from the reference that are deliberate:
 *   - 64-bit indexing everywhere (the reference's `int idx = i*channels`
 *     overflows above 715,827,882 pixels, parallel_to_grayscale.c:9).
 *   - image dimensions are run-time parameters (old/ uses macros from a utils.h
 *     that is missing from the reference).
 *   - uint8 storage variants of the old/ kernels (the reference stores `long`);
 *     the `long` forms are kept too (suffix _long) for 1:1 comparison.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared oracle.c -o liboracle.so -lm
 * (-ffp-contract=off keeps the strict luma strict on FMA hosts; the contracted
 * variant below uses fmaf() explicitly, so it does not depend on compiler flags.)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define ORACLE_API __attribute__((visibility("default")))

/* --- a1: monolithic/src/parallel_to_grayscale.c:5-17 ------------------------
 * lum = (unsigned char)(0.299f*r + 0.587f*g + 0.114f*b), written to bytes 0..2,
 * byte 3 (alpha) untouched.  Strict variant: every product and sum rounded to
 * fp32 (what -O3 -fopenmp produces, microservices/grayscale/c/Makefile:2). */
static inline unsigned char luma_strict(unsigned char r, unsigned char g, unsigned char b)
{
    volatile float pr = 0.299f * (float)r;   /* volatile: forbid re-association/contraction */
    volatile float pg = 0.587f * (float)g;
    volatile float pb = 0.114f * (float)b;
    volatile float t = pr + pg;
    volatile float v = t + pb;
    return (unsigned char)v;
}

/* FMA-contracted variant: what gcc 13 emits for the same expression under
 * monolithic/Makefile:2 (-O3 -march=native -ffast-math) on an FMA host:
 * fma(0.114f,b, fma(0.299f,r, 0.587f*g))  (SURVEY.md appendix A.1). */
static inline unsigned char luma_fma(unsigned char r, unsigned char g, unsigned char b)
{
    volatile float pg = 0.587f * (float)g;
    float t = fmaf(0.299f, (float)r, pg);
    float v = fmaf(0.114f, (float)b, t);
    return (unsigned char)v;
}

ORACLE_API void oracle_gray_inplace(unsigned char *data, long long npixels, int channels, int fma_mode)
{
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < npixels; ++i) {
        unsigned char *p = data + i * (long long)channels;
        unsigned char l = fma_mode ? luma_fma(p[0], p[1], p[2]) : luma_strict(p[0], p[1], p[2]);
        p[0] = l; p[1] = l; p[2] = l;
    }
}

/* 2^24-entry table lut[(r<<16)|(g<<8)|b] of the luma function. */
ORACLE_API void oracle_gray_lut(unsigned char *lut, int fma_mode)
{
    #pragma omp parallel for schedule(static)
    for (int r = 0; r < 256; ++r)
        for (int g = 0; g < 256; ++g)
            for (int b = 0; b < 256; ++b)
                lut[(r << 16) | (g << 8) | b] = fma_mode ? luma_fma(r, g, b) : luma_strict(r, g, b);
}

/* packed variant used by the fused paths: RGB(A) in, one luma byte per pixel out */
ORACLE_API void oracle_rgb_to_luma(const unsigned char *rgb, size_t rgb_pitch, unsigned char *luma,
                                   size_t luma_pitch, int w, int h, int channels, int fma_mode)
{
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const unsigned char *s = rgb + (size_t)y * rgb_pitch;
        unsigned char *d = luma + (size_t)y * luma_pitch;
        for (int x = 0; x < w; ++x, s += channels)
            d[x] = fma_mode ? luma_fma(s[0], s[1], s[2]) : luma_strict(s[0], s[1], s[2]);
    }
}

/* --- a2: monolithic/src/sobel.c:6-25 -----------------------------------------
 * interior pixels only; gx, gy as written at :14-19; mag=(int)sqrtf(gx^2+gy^2)
 * clamped to 255 (:20-21).  Border bytes of dst are never written (:11-12). */
ORACLE_API void oracle_sobel(const unsigned char *src, size_t src_pitch, unsigned char *dst,
                             size_t dst_pitch, int w, int h)
{
    #pragma omp parallel for schedule(static)
    for (int y = 1; y < h - 1; ++y) {
        const unsigned char *a = src + (size_t)(y - 1) * src_pitch;
        const unsigned char *b = a + src_pitch;
        const unsigned char *c = b + src_pitch;
        unsigned char *d = dst + (size_t)y * dst_pitch;
        for (int x = 1; x < w - 1; ++x) {
            int gx = -a[x - 1] - 2 * b[x - 1] - c[x - 1] + a[x + 1] + 2 * b[x + 1] + c[x + 1];
            int gy = a[x - 1] + 2 * a[x] + a[x + 1] - c[x - 1] - 2 * c[x] - c[x + 1];
            int mag = (int)sqrtf((float)(gx * gx + gy * gy));
            if (mag > 255) mag = 255;
            d[x] = (unsigned char)mag;
        }
    }
}

/* Pure-integer form of the same function (no libm): min(255, isqrt(gx^2+gy^2)).
 * Used to prove the sqrtf form has an exact integer characterisation. */
static inline int isqrt_u32(uint32_t v)
{
    uint32_t r = 0, bit = 1u << 30;
    while (bit > v) bit >>= 2;
    while (bit) {
        if (v >= r + bit) { v -= r + bit; r = (r >> 1) + bit; } else r >>= 1;
        bit >>= 2;
    }
    return (int)r;
}

ORACLE_API void oracle_sobel_isqrt(const unsigned char *src, size_t src_pitch, unsigned char *dst,
                                   size_t dst_pitch, int w, int h)
{
    #pragma omp parallel for schedule(static)
    for (int y = 1; y < h - 1; ++y) {
        const unsigned char *a = src + (size_t)(y - 1) * src_pitch;
        const unsigned char *b = a + src_pitch;
        const unsigned char *c = b + src_pitch;
        unsigned char *d = dst + (size_t)y * dst_pitch;
        for (int x = 1; x < w - 1; ++x) {
            int gx = (a[x + 1] + 2 * b[x + 1] + c[x + 1]) - (a[x - 1] + 2 * b[x - 1] + c[x - 1]);
            int gy = (a[x - 1] + 2 * a[x] + a[x + 1]) - (c[x - 1] + 2 * c[x] + c[x + 1]);
            int mag = isqrt_u32((uint32_t)(gx * gx + gy * gy));
            d[x] = (unsigned char)(mag > 255 ? 255 : mag);
        }
    }
}

/* --- a3..a5: monolithic/src/main_with_sobel.c:51-74 ---------------------------
 * one pass of the driver's loop: gray in place (:54), extract channel 0 (:57-59),
 * sobel gray->edge (:62), replicate edge into R,G,B (:65-73; alpha untouched).
 * `gray` and `edge` are caller scratch planes of w*h bytes, exactly like the
 * driver's malloc'ed buffers (:36-37); edge's border keeps whatever it held. */
ORACLE_API void oracle_gray_sobel_pass(unsigned char *img, int w, int h, int channels,
                                       unsigned char *gray, unsigned char *edge, int fma_mode)
{
    long long np = (long long)w * h;
    oracle_gray_inplace(img, np, channels, fma_mode);
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < np; ++i) gray[i] = img[i * channels];
    oracle_sobel(gray, (size_t)w, edge, (size_t)w, w, h);
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < np; ++i) {
        unsigned char e = edge[i];
        unsigned char *p = img + i * channels;
        p[0] = e; p[1] = e; p[2] = e;
    }
}

/* The driver's two glue loops on their own (main_with_sobel.c:57-59 and :65-73), so that
 * bench.py --impl reference can time the compiled reference functions (oracle/_ref) around them. */
ORACLE_API void oracle_extract_plane(const unsigned char *img, long long npixels, int channels, unsigned char *gray)
{
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < npixels; ++i) gray[i] = img[i * channels];
}

ORACLE_API void oracle_replicate_plane(const unsigned char *edge, long long npixels, int channels, unsigned char *img)
{
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < npixels; ++i) {
        unsigned char e = edge[i];
        unsigned char *p = img + i * channels;
        p[0] = e; p[1] = e; p[2] = e;
    }
}

/* Fused statement of the same pass without the RGB round trip: RGB(A) in, edge
 * plane out (interior only).  Equal to sobel(luma(rgb)) by construction. */
ORACLE_API void oracle_rgb_to_edge(const unsigned char *rgb, size_t rgb_pitch, int channels,
                                   unsigned char *edge, size_t edge_pitch, int w, int h,
                                   unsigned char *luma_scratch, int fma_mode)
{
    oracle_rgb_to_luma(rgb, rgb_pitch, luma_scratch, (size_t)w, w, h, channels, fma_mode);
    oracle_sobel(luma_scratch, (size_t)w, edge, edge_pitch, w, h);
}

/* --- a6: old/parallel_convolution.c:8-24 --------------------------------------
 * out[r][c][p] += sum_{kr,kc} img[r+kr][c+kc][p]*kernel[kr][kc]; out /= NORM.
 * Input pre-padded (rows+K-1 x cols+K-1), top-left anchored, C truncating
 * division, no clamp, accumulates onto the prior contents of out. */
ORACLE_API void oracle_convolution_long(const long *img, int rows, int cols, int K,
                                        const long *kernel, long norm, long *out)
{
    int pc = cols + K - 1;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            for (int p = 0; p < 3; ++p) {
                long acc = out[((size_t)r * cols + c) * 3 + p];
                for (int kr = 0; kr < K; ++kr)
                    for (int kc = 0; kc < K; ++kc)
                        acc += img[((size_t)(r + kr) * pc + (c + kc)) * 3 + p] * kernel[kr * K + kc];
                out[((size_t)r * cols + c) * 3 + p] = acc / norm;
            }
}

/* Same arithmetic on uint8 interleaved RGB storage with int32 taps.  The result
 * is returned as int32 (no clamp, like the reference); out is overwritten
 * (equivalent to the reference with a zeroed output).  in_pitch in bytes. */
ORACLE_API void oracle_convolution_u8(const unsigned char *img, size_t in_pitch, int rows, int cols,
                                      int K, const int *kernel, int norm, int *out)
{
    #pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            for (int p = 0; p < 3; ++p) {
                long long acc = 0;
                for (int kr = 0; kr < K; ++kr) {
                    const unsigned char *row = img + (size_t)(r + kr) * in_pitch + (size_t)c * 3 + p;
                    for (int kc = 0; kc < K; ++kc)
                        acc += (long long)row[kc * 3] * kernel[kr * K + kc];
                }
                out[((size_t)r * cols + c) * 3 + p] = (int)(acc / norm);
            }
}

/* --- a7: old/parallel_avg_pixel.c:5-42 ----------------------------------------
 * every channel is summed three times (inner `pixel` loop, :19-23), added onto
 * avgs (:28-30) and divided by rows*cols*3 (:8,:40): floor-toward-zero of
 * (avgs_in + 3*sum_c) / (rows*cols*3). */
ORACLE_API void oracle_avg_pixel_long(const long *img, int rows, int cols, long *avgs)
{
    long s0 = 0, s1 = 0, s2 = 0;
    long count = (long)rows * cols * 3;
    #pragma omp parallel for reduction(+:s0, s1, s2) schedule(static)
    for (long long i = 0; i < (long long)rows * cols; ++i) {
        s0 += 3 * img[i * 3 + 0]; s1 += 3 * img[i * 3 + 1]; s2 += 3 * img[i * 3 + 2];
    }
    avgs[0] = (avgs[0] + s0) / count; avgs[1] = (avgs[1] + s1) / count; avgs[2] = (avgs[2] + s2) / count;
}

ORACLE_API void oracle_avg_pixel_u8(const unsigned char *img, size_t pitch, int rows, int cols,
                                    long long *avgs)
{
    long long s0 = 0, s1 = 0, s2 = 0;
    long long count = (long long)rows * cols * 3;
    #pragma omp parallel for reduction(+:s0, s1, s2) schedule(static)
    for (int r = 0; r < rows; ++r) {
        const unsigned char *p = img + (size_t)r * pitch;
        for (int c = 0; c < cols; ++c, p += 3) { s0 += 3 * p[0]; s1 += 3 * p[1]; s2 += 3 * p[2]; }
    }
    avgs[0] = (avgs[0] + s0) / count; avgs[1] = (avgs[1] + s1) / count; avgs[2] = (avgs[2] + s2) / count;
}

/* --- a8: old/parallel_to_grayscale.c:7-38 --------------------------------------
 * v=(r+g+b)/3 to all three channels (:15-22); min/max start at 256 / -1 (:9-10). */
ORACLE_API void oracle_gray_avg3_long(const long *img, int rows, int cols, long *gray, long *min_max)
{
    int mn = 256, mx = -1;
    #pragma omp parallel for reduction(min:mn) reduction(max:mx) schedule(static)
    for (long long i = 0; i < (long long)rows * cols; ++i) {
        int sum = (int)(img[i * 3] + img[i * 3 + 1] + img[i * 3 + 2]);
        int v = sum / 3;
        gray[i * 3] = v; gray[i * 3 + 1] = v; gray[i * 3 + 2] = v;
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
    min_max[0] = mn; min_max[1] = mx;
}

ORACLE_API void oracle_gray_avg3_u8(const unsigned char *img, size_t in_pitch, unsigned char *gray,
                                    size_t out_pitch, int rows, int cols, int *min_max)
{
    int mn = 256, mx = -1;
    #pragma omp parallel for reduction(min:mn) reduction(max:mx) schedule(static)
    for (int r = 0; r < rows; ++r) {
        const unsigned char *s = img + (size_t)r * in_pitch;
        unsigned char *d = gray + (size_t)r * out_pitch;
        for (int c = 0; c < cols; ++c) {
            int v = (s[c * 3] + s[c * 3 + 1] + s[c * 3 + 2]) / 3;
            d[c * 3] = d[c * 3 + 1] = d[c * 3 + 2] = (unsigned char)v;
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
    }
    min_max[0] = mn; min_max[1] = mx;
}

/* --- synthetic input: counter-based hash, identical to the
 * device generator (csrc/synth.cuh) and to tests/common.py. ------------------- */
static inline uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

ORACLE_API void oracle_synth_fill(unsigned char *dst, unsigned long long nbytes, unsigned int seed,
                                  unsigned long long byte_offset)
{
    #pragma omp parallel for schedule(static)
    for (unsigned long long i = 0; i < nbytes; ++i) {
        unsigned long long gi = byte_offset + i;
        uint32_t w = mix32(seed + (uint32_t)(gi >> 2) * 0x9E3779B9u);
        dst[i] = (unsigned char)(w >> (8 * (gi & 3)));
    }
}

/* --- harness helpers (no pixel math) ------------------------------------------
 * The OpenMP thread count the runtime will really use (bench.py prints this, not os.cpu_count()). */
#include <omp.h>
ORACLE_API int oracle_omp_max_threads(void) { return omp_get_max_threads(); }

/* The reference's old/ output parameters are `long ***` tables (old/parallel_convolution.c:8,
 * old/parallel_to_grayscale.c:7); building one with a million ctypes casts is too slow for the timed
 * CPU baseline, so it is built here over a dense rows x cols x 3 `long` array. */
ORACLE_API void *oracle_table3_alloc(long *base, int rows, int cols)
{
    long ***top = (long ***)malloc((size_t)rows * sizeof(long **) + (size_t)rows * cols * sizeof(long *));
    if (!top) return NULL;
    long **cells = (long **)(top + rows);
    for (int r = 0; r < rows; ++r) {
        top[r] = cells + (size_t)r * cols;
        for (int c = 0; c < cols; ++c) top[r][c] = base + ((size_t)r * cols + c) * 3;
    }
    return top;
}
ORACLE_API void oracle_table3_free(void *t) { free(t); }

/* --- f2: the JPEG back end (stb_image v2.30 as vendored by the reference, monolithic/include/stb_image.h) -----------
 * TEST INFRASTRUCTURE: a scalar restatement of what stb does AFTER entropy decoding, used (a) on the CPU to check the
 * product's host-side entropy decoder without a GPU and (b) as a second statement of the arithmetic next to the real
 * checker, the reference's own stb_image.h compiled as it lies (oracle/tools/stb_decode.c -> oracle/_ref/bin/stb_decode).
 *   dequantise  stb_image.h:2227,2245,3073-3078 (short arithmetic)     IDCT      stb_image.h:2430-2524
 *   resampling  stb_image.h:3464-3527, row pairing :3907-3941           colour    stb_image.h:3659-3683
 * comp[k]: coeff [bh][bw][64] natural order, dq[64]; planes are produced internally.  out: w*h*(ncomp==3?3:1) bytes. */
#define OJ_F2F(x) ((int)(((x) * 4096 + 0.5)))
#define OJ_IDCT_1D(s0,s1,s2,s3,s4,s5,s6,s7) \
   int t0,t1,t2,t3,p1,p2,p3,p4,p5,x0,x1,x2,x3; \
   p2 = s2; p3 = s6; p1 = (p2+p3) * OJ_F2F(0.5411961f); t2 = p1 + p3*OJ_F2F(-1.847759065f); t3 = p1 + p2*OJ_F2F(0.765366865f); \
   p2 = s0; p3 = s4; t0 = (p2+p3)*4096; t1 = (p2-p3)*4096; x0 = t0+t3; x3 = t0-t3; x1 = t1+t2; x2 = t1-t2; \
   t0 = s7; t1 = s5; t2 = s3; t3 = s1; p3 = t0+t2; p4 = t1+t3; p1 = t0+t3; p2 = t1+t2; p5 = (p3+p4)*OJ_F2F(1.175875602f); \
   t0 = t0*OJ_F2F(0.298631336f); t1 = t1*OJ_F2F(2.053119869f); t2 = t2*OJ_F2F(3.072711026f); t3 = t3*OJ_F2F(1.501321110f); \
   p1 = p5 + p1*OJ_F2F(-0.899976223f); p2 = p5 + p2*OJ_F2F(-2.562915447f); p3 = p3*OJ_F2F(-1.961570560f); p4 = p4*OJ_F2F(-0.390180644f); \
   t3 += p1+p4; t2 += p2+p3; t1 += p2+p4; t0 += p1+p3;
static unsigned char oj_clamp(int x) { return (unsigned char)(x < 0 ? 0 : x > 255 ? 255 : x); }
static void oj_idct(unsigned char *out, int stride, const short *in, const unsigned short *dq)
{
    short d[64];
    int v[64];
    for (int i = 0; i < 64; ++i) d[i] = (short)(in[i] * dq[i]);
    for (int i = 0; i < 8; ++i) {
        OJ_IDCT_1D(d[i], d[8+i], d[16+i], d[24+i], d[32+i], d[40+i], d[48+i], d[56+i])
        x0 += 512; x1 += 512; x2 += 512; x3 += 512;
        v[i] = (x0+t3) >> 10; v[56+i] = (x0-t3) >> 10; v[8+i] = (x1+t2) >> 10; v[48+i] = (x1-t2) >> 10;
        v[16+i] = (x2+t1) >> 10; v[40+i] = (x2-t1) >> 10; v[24+i] = (x3+t0) >> 10; v[32+i] = (x3-t0) >> 10;
    }
    for (int i = 0; i < 8; ++i) {
        unsigned char *o = out + (size_t)i * stride;
        OJ_IDCT_1D(v[8*i], v[8*i+1], v[8*i+2], v[8*i+3], v[8*i+4], v[8*i+5], v[8*i+6], v[8*i+7])
        x0 += 65536 + (128<<17); x1 += 65536 + (128<<17); x2 += 65536 + (128<<17); x3 += 65536 + (128<<17);
        o[0] = oj_clamp((x0+t3) >> 17); o[7] = oj_clamp((x0-t3) >> 17); o[1] = oj_clamp((x1+t2) >> 17); o[6] = oj_clamp((x1-t2) >> 17);
        o[2] = oj_clamp((x2+t1) >> 17); o[5] = oj_clamp((x2-t1) >> 17); o[3] = oj_clamp((x3+t0) >> 17); o[4] = oj_clamp((x3-t0) >> 17);
    }
}
/* one output row of component k, resampled to full width like stb's row functions; returns the row to read from */
static const unsigned char *oj_resample(unsigned char *buf, const unsigned char *near, const unsigned char *far, int w, int hs, int vs)
{
    if (hs == 1 && vs == 1) return near;
    if (hs == 1) { for (int i = 0; i < w; ++i) buf[i] = (unsigned char)((3*near[i] + far[i] + 2) >> 2); return buf; }
    if (vs == 1) {
        if (w == 1) { buf[0] = buf[1] = near[0]; return buf; }
        buf[0] = near[0]; buf[1] = (unsigned char)((near[0]*3 + near[1] + 2) >> 2);
        int i;
        for (i = 1; i < w-1; ++i) { int n = 3*near[i]+2; buf[i*2] = (unsigned char)((n+near[i-1]) >> 2); buf[i*2+1] = (unsigned char)((n+near[i+1]) >> 2); }
        buf[i*2] = (unsigned char)((near[w-2]*3 + near[w-1] + 2) >> 2); buf[i*2+1] = near[w-1];
        return buf;
    }
    if (w == 1) { buf[0] = buf[1] = (unsigned char)((3*near[0] + far[0] + 2) >> 2); return buf; }
    int t0, t1 = 3*near[0] + far[0];
    buf[0] = (unsigned char)((t1+2) >> 2);
    for (int i = 1; i < w; ++i) { t0 = t1; t1 = 3*near[i]+far[i]; buf[i*2-1] = (unsigned char)((3*t0 + t1 + 8) >> 4); buf[i*2] = (unsigned char)((3*t1 + t0 + 8) >> 4); }
    buf[w*2-1] = (unsigned char)((t1+2) >> 2);
    return buf;
}
#define OJ_FIX(x) (((int)((x) * 4096.0f + 0.5f)) << 8)
ORACLE_API int oracle_jpeg_backend(int width, int height, int ncomp, int hmax, int vmax, int ycbcr,
                                   const short *const *coeff, const unsigned short *const *dq, const int *ch, const int *cv,
                                   const int *bw, const int *bh, const int *cy, unsigned char *out)
{
    unsigned char *plane[3] = {0, 0, 0}, *line[3] = {0, 0, 0};
    if (ncomp != 1 && ncomp != 3) return -1;
    for (int k = 0; k < ncomp; ++k) {
        plane[k] = (unsigned char *)malloc((size_t)bw[k] * bh[k] * 64);
        line[k] = (unsigned char *)malloc((size_t)width + 8);
        for (int by = 0; by < bh[k]; ++by)
            for (int bx = 0; bx < bw[k]; ++bx)
                oj_idct(plane[k] + ((size_t)by * 8) * bw[k] * 8 + bx * 8, bw[k] * 8, coeff[k] + ((size_t)by * bw[k] + bx) * 64, dq[k]);
    }
    const int n = ncomp == 3 ? 3 : 1;
    int ystep[3], ypos[3];
    const unsigned char *l0[3], *l1[3];
    for (int k = 0; k < ncomp; ++k) { ystep[k] = (vmax / cv[k]) >> 1; ypos[k] = 0; l0[k] = l1[k] = plane[k]; }
    for (int j = 0; j < height; ++j) {
        const unsigned char *row[3];
        for (int k = 0; k < ncomp; ++k) {
            const int hs = hmax / ch[k], vs = vmax / cv[k], y_bot = ystep[k] >= (vs >> 1);
            row[k] = oj_resample(line[k], y_bot ? l1[k] : l0[k], y_bot ? l0[k] : l1[k], (width + hs - 1) / hs, hs, vs);
            if (++ystep[k] >= vs) { ystep[k] = 0; l0[k] = l1[k]; if (++ypos[k] < cy[k]) l1[k] += bw[k] * 8; }
        }
        unsigned char *o = out + (size_t)j * width * n;
        for (int i = 0; i < width; ++i, o += n) {
            if (ncomp == 1) { o[0] = row[0][i]; continue; }
            if (!ycbcr) { o[0] = row[0][i]; o[1] = row[1][i]; o[2] = row[2][i]; continue; }
            int y_fixed = (row[0][i] << 20) + (1<<19), cr = row[2][i] - 128, cb = row[1][i] - 128;
            int r = y_fixed + cr * OJ_FIX(1.40200f);
            int g = y_fixed + (cr * -OJ_FIX(0.71414f)) + ((cb * -OJ_FIX(0.34414f)) & 0xffff0000);
            int b = y_fixed + cb * OJ_FIX(1.77200f);
            o[0] = oj_clamp(r >> 20); o[1] = oj_clamp(g >> 20); o[2] = oj_clamp(b >> 20);
        }
    }
    for (int k = 0; k < ncomp; ++k) { free(plane[k]); free(line[k]); }
    return 0;
}
