/* tests/adapters/old_long_adapters.c -- TEST-ONLY adapters: the three functions of the reference's old/
 * directory, with the reference's own `long`-array signatures (old/parallel_convolution.c:8, old/parallel_avg_pixel.c:5,
 * old/parallel_to_grayscale.c:7), implemented on libompimg_b200's C ABI.  Compiled by tests/ against the same stub
 * utils.h (-DDIM_ROW ... like oracle/Makefile's OLD_RULE) so that a test can call the reference object and this
 * object with the SAME arguments and compare the `long` arrays.  Not part of the product: the product API stores
 * pixels as bytes; these only narrow long -> uint8 on the way in (pixels are 0..255 throughout the reference)
 * and widen on the way out.  No arithmetic happens here. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "utils.h"
#include "ompimg_b200.h"

static void die(const char *what) {
    fprintf(stderr, "old_long_adapters: %s: %s\n", what, ompimg_last_error());
    abort();
}
static unsigned char *narrow(const long *src, size_t n) {
    unsigned char *d = (unsigned char *)malloc(n ? n : 1);
    if (!d) die("malloc");
    for (size_t i = 0; i < n; ++i) {
        if (src[i] < 0 || src[i] > 255) { fprintf(stderr, "old_long_adapters: pixel value %ld outside 0..255\n", src[i]); abort(); }
        d[i] = (unsigned char)src[i];
    }
    return d;
}

void parallel_convolution(long img[DIM_ROW + PAD][DIM_COL + PAD][DIM_RGB], long kernel[DIM_KERNEL][DIM_KERNEL], long ***convolved_img) {
    const size_t n_in = (size_t)(DIM_ROW + PAD) * (DIM_COL + PAD) * DIM_RGB, n_out = (size_t)DIM_ROW * DIM_COL * DIM_RGB;
    unsigned char *in = narrow(&img[0][0][0], n_in);
    int32_t taps[DIM_KERNEL * DIM_KERNEL];
    for (int i = 0; i < DIM_KERNEL; ++i)
        for (int j = 0; j < DIM_KERNEL; ++j) taps[i * DIM_KERNEL + j] = (int32_t)kernel[i][j];
    int32_t *out = (int32_t *)malloc(n_out * sizeof(int32_t));
    if (!out) die("malloc");
    for (int r = 0; r < DIM_ROW; ++r)                      /* the reference accumulates onto what the caller's table holds (:17) */
        for (int c = 0; c < DIM_COL; ++c)
            for (int p = 0; p < DIM_RGB; ++p) out[((size_t)r * DIM_COL + c) * DIM_RGB + p] = (int32_t)convolved_img[r][c][p];
    if (ompimg_conv_i32_accum_host(in, DIM_ROW, DIM_COL, DIM_KERNEL, taps, GBLUR_NORM, out) != OMPIMG_OK) die("ompimg_conv_i32_accum_host");
    for (int r = 0; r < DIM_ROW; ++r)
        for (int c = 0; c < DIM_COL; ++c)
            for (int p = 0; p < DIM_RGB; ++p) convolved_img[r][c][p] = out[((size_t)r * DIM_COL + c) * DIM_RGB + p];
    free(out);
    free(in);
}

void parallel_avg_pixel(long img[DIM_ROW][DIM_COL][DIM_RGB], long *avgs) {
    unsigned char *in = narrow(&img[0][0][0], (size_t)DIM_ROW * DIM_COL * DIM_RGB);
    long long a[3] = {avgs[0], avgs[1], avgs[2]};
    if (ompimg_avg_pixel_host(in, DIM_ROW, DIM_COL, a) != OMPIMG_OK) die("ompimg_avg_pixel_host");
    for (int c = 0; c < 3; ++c) avgs[c] = (long)a[c];
    free(in);
}

void parallel_to_grayscale(long img[DIM_ROW][DIM_COL][DIM_RGB], long ***grayscale_img, long *min_max_gray) {
    const size_t n = (size_t)DIM_ROW * DIM_COL * DIM_RGB;
    unsigned char *in = narrow(&img[0][0][0], n);
    unsigned char *out = (unsigned char *)malloc(n);
    int mm[2];
    if (!out) die("malloc");
    if (ompimg_gray_avg3_host(in, out, DIM_ROW, DIM_COL, mm) != OMPIMG_OK) die("ompimg_gray_avg3_host");
    for (int r = 0; r < DIM_ROW; ++r)
        for (int c = 0; c < DIM_COL; ++c)
            for (int p = 0; p < DIM_RGB; ++p) grayscale_img[r][c][p] = out[((size_t)r * DIM_COL + c) * DIM_RGB + p];
    min_max_gray[0] = mm[0];
    min_max_gray[1] = mm[1];
    free(out);
    free(in);
}
