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
