/*
 * ompimg_b200.h -- C ABI of libompimg_b200.so, the B200 (sm_100a) implementation of the
 * per-pixel hot path of PedemonteGiacomo/OpenMp-Parallel-Computing.
 *
 * Plain C: pointers, sizes and ints only.  Every entry point cites the reference code it
 * replaces (paths relative to the reference repository root).  There is no CPU fallback:
 * when no CUDA device is usable every int-returning call fails with OMPIMG_ERR_CUDA and the
 * two frozen void functions print one line to stderr and abort().
 *
 * Conventions
 *   - "_dev" entry points take DEVICE pointers and enqueue work on one of the library's
 *     streams (stream index 0..OMPIMG_NUM_STREAMS-1); they return without synchronising.
 *   - "_host" entry points take HOST pointers (pageable or pinned), are synchronous, and move
 *     the data through pinned staging with H2D / kernel / D2H overlapped on several streams.
 *   - pitches and strides are in bytes.  Fast paths need 16-byte aligned base pointers and
 *     pitches that are multiples of 16; anything else takes a slower generic kernel with the
 *     same results.
 *   - luma mode: OMPIMG_LUMA_STRICT reproduces the reference built with `-O3 -fopenmp`
 *     (microservices/grayscale/c/Makefile:2; pinned by images/test_gray.png),
 *     OMPIMG_LUMA_FMA reproduces `-O3 -march=native -ffast-math` on an FMA host
 *     (monolithic/Makefile:2).  Default strict; env OMPIMG_LUMA=strict|fma overrides.
 *   - return value: 0 on success, a negative OMPIMG_ERR_* otherwise; ompimg_last_error()
 *     gives the text.
 */
#ifndef OMPIMG_B200_H
#define OMPIMG_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define OMPIMG_API __attribute__((visibility("default")))
#else
#define OMPIMG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define OMPIMG_OK            0
#define OMPIMG_ERR_CUDA     -1   /* CUDA runtime/driver error or no device */
#define OMPIMG_ERR_ARG      -2   /* invalid argument */
#define OMPIMG_ERR_NOMEM    -3

#define OMPIMG_LUMA_STRICT   0
#define OMPIMG_LUMA_FMA      1

#define OMPIMG_NUM_STREAMS   8

/* ---- frozen prototypes -------------------------------------------------------------------
 * monolithic/include/parallel_to_grayscale.h:4  (implementation replaced:
 * monolithic/src/parallel_to_grayscale.c:5-17) and monolithic/include/sobel.h:3-5
 * (monolithic/src/sobel.c:6-25).  Same names, same arguments, same in-place / border-untouched
 * semantics, synchronous.  Errors abort the process (non-zero exit), which the Flask and
 * RabbitMQ workers map to HTTP 500 / CalledProcessError. */
OMPIMG_API void convert_to_grayscale(unsigned char *data, int width, int height, int channels);
OMPIMG_API void sobel_edge(const unsigned char *src, unsigned char *dst, int width, int height);

/* ---- runtime ------------------------------------------------------------------------------ */
OMPIMG_API int         ompimg_init(int device);            /* bind this process/thread group to a device; lazy otherwise (env OMPIMG_DEVICE) */
OMPIMG_API int         ompimg_device(void);                /* bound device or -1 */
OMPIMG_API int         ompimg_device_count(void);
OMPIMG_API int         ompimg_sm_count(void);
OMPIMG_API const char *ompimg_last_error(void);
OMPIMG_API const char *ompimg_version(void);
OMPIMG_API int         ompimg_set_luma_mode(int mode);     /* OMPIMG_LUMA_* */
OMPIMG_API int         ompimg_get_luma_mode(void);
OMPIMG_API int         ompimg_sqrt_is_exact(void);   /* 1: the device passed the exhaustive sqrt check at init (no +0.5 guard in the Sobel kernels), 0: guard in use, -1: not initialised */
OMPIMG_API long long   ompimg_launch_count(void);          /* kernels launched by this library so far */
OMPIMG_API int         ompimg_shutdown(void);

OMPIMG_API void *ompimg_dev_alloc(size_t bytes);           /* cudaMalloc, NULL on failure */
OMPIMG_API int   ompimg_dev_free(void *p);
OMPIMG_API void *ompimg_host_alloc(size_t bytes);          /* pinned host memory */
OMPIMG_API int   ompimg_host_free(void *p);
OMPIMG_API void *ompimg_host_realloc_sized(void *p, size_t old_bytes, size_t new_bytes);   /* for STBI_REALLOC_SIZED (stb_image.h:673-684) */
OMPIMG_API int   ompimg_host_register(void *p, size_t bytes);   /* pin caller-owned pageable memory */
OMPIMG_API int   ompimg_host_unregister(void *p);
OMPIMG_API int   ompimg_memset_dev(void *dst, int value, size_t bytes, int stream);
OMPIMG_API int   ompimg_copy_h2d(void *dst, const void *src, size_t bytes, int stream);
OMPIMG_API int   ompimg_copy_d2h(void *dst, const void *src, size_t bytes, int stream);
OMPIMG_API int   ompimg_copy_d2d(void *dst, const void *src, size_t bytes, int stream);
OMPIMG_API int   ompimg_copy2d_h2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, int stream);
OMPIMG_API int   ompimg_copy2d_d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, int stream);
OMPIMG_API int   ompimg_stream_sync(int stream);           /* stream < 0: all streams */
OMPIMG_API int   ompimg_stream_wait_stream(int waiter, int signaller);   /* event edge between two library streams */

/* device-side timers (CUDA events recorded on a library stream) */
OMPIMG_API int   ompimg_timer_create(void);                /* returns timer id >= 0 */
OMPIMG_API int   ompimg_timer_start(int timer, int stream);
OMPIMG_API int   ompimg_timer_stop(int timer, int stream);
OMPIMG_API int   ompimg_timer_elapsed_ms(int timer, float *ms);  /* synchronises on the stop event */
OMPIMG_API int   ompimg_timer_destroy(int timer);

/* ---- a1: grayscale (monolithic/src/parallel_to_grayscale.c:5-17) ------------------------------
 * lum = (unsigned char)(0.299f*r + 0.587f*g + 0.114f*b) written to bytes 0..2 of every pixel,
 * byte 3 (alpha) untouched; channels in {3,4} (1 and 2 are UB in the reference: rejected). */
OMPIMG_API int ompimg_gray_inplace_dev(uint8_t *img, size_t pitch, int width, int height, int channels,
                            int nframes, size_t frame_stride, int stream);
/* packed variant used by fused pipelines: one luma byte per pixel */
OMPIMG_API int ompimg_rgb_to_luma_dev(const uint8_t *rgb, size_t rgb_pitch, int channels,
                           uint8_t *luma, size_t luma_pitch, int width, int height,
                           int nframes, size_t rgb_stride, size_t luma_stride, int stream);
OMPIMG_API int ompimg_gray_inplace_host(uint8_t *data, int width, int height, int channels);

/* ---- a2: Sobel (monolithic/src/sobel.c:6-25) ---------------------------------------------------
 * dst[y][x] = min(255, (int)sqrtf(gx*gx+gy*gy)) for 1<=y<=h-2, 1<=x<=w-2; every other byte of
 * dst is left untouched.  src and dst must not overlap. */
OMPIMG_API int ompimg_sobel_dev(const uint8_t *src, size_t src_pitch, uint8_t *dst, size_t dst_pitch,
                     int width, int height, int nframes, size_t src_stride, size_t dst_stride, int stream);
OMPIMG_API int ompimg_sobel_host(const uint8_t *src, uint8_t *dst, int width, int height);

/* ---- a3..a5: the gray -> extract -> Sobel -> replicate pass (monolithic/src/main_with_sobel.c:51-74)
 * fused: RGB(A) in, edge plane out; the luma plane never goes to HBM.  Interior only, like a2. */
OMPIMG_API int ompimg_rgb_to_edge_dev(const uint8_t *rgb, size_t rgb_pitch, int channels,
                           uint8_t *edge, size_t edge_pitch, int width, int height,
                           int nframes, size_t rgb_stride, size_t edge_stride, int stream);
/* one row band of a larger image (multi-GPU / multi-process row-band mode): `rgb` holds `band_rows`
 * rows (channels 1 = luma plane, 3/4 = RGB(A)); row_above / row_below are DEVICE-readable
 * addresses (local, peer-mapped or NULL at the image's top/bottom) of the source rows adjacent to
 * the band.  Every band row that has both neighbours is written. */
OMPIMG_API int ompimg_rgb_to_edge_band_dev(const uint8_t *rgb, size_t rgb_pitch, int channels, uint8_t *edge, size_t edge_pitch,
                                int width, int band_rows, const uint8_t *row_above, const uint8_t *row_below, int stream);
/* a4 (main_with_sobel.c:65-73): img[i*ch+0..2] = plane[i]; alpha untouched */
OMPIMG_API int ompimg_replicate_dev(const uint8_t *plane, size_t plane_pitch, uint8_t *img, size_t img_pitch,
                         int channels, int width, int height, int nframes,
                         size_t plane_stride, size_t img_stride, int stream);
/* one iteration of the driver loop (main_with_sobel.c:51-74), in place on img.  edge is the driver's persistent w*h scratch
 * plane (main_with_sobel.c:37): its border bytes are what the image border becomes, exactly as in the reference; edge == NULL
 * means a zero border.  3-channel images whose rows are 16-byte friendly take the fused path: a small preparation kernel and
 * ONE kernel that reads RGB and writes the replicated edge value back (6 bytes/pixel; the edge plane's interior is then
 * NOT written -- the plane never exists).  Other layouts run the fused gray->Sobel kernel into `edge` and a replicate
 * kernel (edge must then be given).  The host form always returns the edge plane in `edge` when it is not NULL. */
OMPIMG_API int ompimg_gray_sobel_pass_dev(uint8_t *img, size_t img_pitch, int channels, uint8_t *edge, size_t edge_pitch,
                               int width, int height, int nframes, size_t img_stride, size_t edge_stride, int stream);
/* the driver's whole loop, `for (p = 0; p < passes; p++)` around main_with_sobel.c:51-74, in place on img: on the fused path every
 * pass but the last also writes what its successor needs (the rows and pixels either side of the kernel's band and strip seams), so
 * only the first pass runs the preparation kernel: passes + 1 launches, 6 bytes/pixel each. */
OMPIMG_API int ompimg_gray_sobel_passes_dev(uint8_t *img, size_t img_pitch, int channels, uint8_t *edge, size_t edge_pitch,
                                 int width, int height, int nframes, size_t img_stride, size_t edge_stride, int passes, int stream);
OMPIMG_API int ompimg_gray_sobel_pass_host(uint8_t *img, int width, int height, int channels, uint8_t *edge);

/* batch of frames in host memory -> edge planes in host memory, H2D/kernel/D2H overlapped on
 * OMPIMG_NUM_STREAMS streams of the bound device.  frames are `nframes` dense RGB(A) images
 * `rgb_stride` bytes apart; edges `edge_stride` bytes apart (dense width*height planes).
 * The border bytes of every edge plane are written as 0. */
OMPIMG_API int ompimg_rgb_to_edge_batch_host(const uint8_t *rgb, size_t rgb_stride, int channels,
                                  uint8_t *edge, size_t edge_stride, int width, int height, int nframes);

/* ---- a6: K x K convolution (old/parallel_convolution.c:8-24) -----------------------------------
 * out[r][c][p] = (sum_{kr,kc} in[r+kr][c+kc][p] * taps[kr*K+kc]) / norm  (C truncating division,
 * no clamp).  Input is pre-padded: (rows+K-1) x (cols+K-1) interleaved uint8 RGB, top-left
 * anchored exactly like the reference.  Output int32 RGB (rows x cols x 3, dense) or, with
 * ompimg_conv_u8_dev, uint8.  The reference does not clamp (old/parallel_convolution.c:20), and neither does the
 * uint8 form: it stores (uint8_t)value, i.e. the low 8 bits of the exact int32 result -- what a C caller gets by
 * narrowing the reference's `long`.  Box/Gaussian taps with norm = sum(taps) never leave 0..255; for taps that can
 * (negative taps, sum(taps) > norm) use ompimg_conv_i32_dev unless that wrap is what you want.  K odd or even, 1..15.
 * in_pitch >= (cols+K-1)*3 and out_pitch >= cols*3 elements of the output type, else OMPIMG_ERR_ARG.
 * Layout: every input row, the last one included, must be readable over its whole pitch (a pitched allocation is rows x pitch).
 * Rank-1 taps (box, binomial, Gaussian; K 3/5/7/9) on a 16-byte aligned base and pitch take the tensor-map kernel (fastest); the
 * reference's dense (cols+K-1)*3 pitch takes a kernel that needs 4-byte alignment only; anything else a generic kernel. */
OMPIMG_API int ompimg_conv_i32_dev(const uint8_t *in, size_t in_pitch, int rows, int cols, int K,
                        const int32_t *taps_host, int32_t norm, int32_t *out, size_t out_pitch, int stream);
OMPIMG_API int ompimg_conv_u8_dev(const uint8_t *in, size_t in_pitch, int rows, int cols, int K,
                       const int32_t *taps_host, int32_t norm, uint8_t *out, size_t out_pitch, int stream);
OMPIMG_API int ompimg_conv_i32_host(const uint8_t *in_padded, int rows, int cols, int K,
                         const int32_t *taps, int32_t norm, int32_t *out);
/* the reference's exact statement: it ACCUMULATES the tap sums onto the caller's array and then divides
 * (old/parallel_convolution.c:17,20), so out = (out_on_entry + sum) / norm.  ompimg_conv_i32_host is the case every
 * caller of the reference sets up (a zeroed array). */
OMPIMG_API int ompimg_conv_i32_accum_host(const uint8_t *in_padded, int rows, int cols, int K,
                               const int32_t *taps, int32_t norm, int32_t *out_inout);

/* ---- a7: average pixel (old/parallel_avg_pixel.c:5-42) -------------------------------------------
 * sums[c] = sum of channel c over rows x cols RGB pixels (exact, 64-bit).  The reference's
 * result is (avgs_in[c] + 3*sums[c]) / (rows*cols*3) with C division; ompimg_avg_pixel_host
 * applies that formula, the _dev form returns the raw sums (so that multi-GPU callers can add
 * partial sums before dividing). */
OMPIMG_API int ompimg_channel_sums_dev(const uint8_t *rgb, size_t pitch, int rows, int cols,
                            unsigned long long *sums3_dev, int stream);   /* adds onto sums3_dev[0..2] */
OMPIMG_API int ompimg_avg_pixel_host(const uint8_t *rgb, int rows, int cols, long long *avgs3_inout);

/* ---- a8: legacy grayscale + min/max (old/parallel_to_grayscale.c:7-38) ---------------------------
 * v = (r+g+b)/3 written to all three channels of out; min_max = {min v, max v} (start 256 / -1). */
OMPIMG_API int ompimg_gray_avg3_dev(const uint8_t *rgb, size_t in_pitch, uint8_t *out, size_t out_pitch,
                         int rows, int cols, int *min_max_dev, int stream);  /* min_max_dev pre-set to {256,-1} */
OMPIMG_API int ompimg_gray_avg3_host(const uint8_t *rgb, uint8_t *out, int rows, int cols, int *min_max);

/* ---- synthetic input: byte i of the stream = mix32(seed + (i>>2)*0x9E3779B9) >> 8*(i&3) */
OMPIMG_API int ompimg_synth_fill_dev(uint8_t *dst, unsigned long long nbytes, unsigned int seed,
                          unsigned long long byte_offset, int stream);

/* checksum helper: adds the sum of the bytes of a width x rows region (pitch bytes between rows) onto *sum_dev */
OMPIMG_API int ompimg_byte_sum_dev(const uint8_t *src, size_t pitch, int width, int rows, unsigned long long *sum_dev, int stream);

/* ---- self tests used by tests/ (device arithmetic against closed forms) --------------------------- */
/* lut[(r<<16)|(g<<8)|b] = luma(r,g,b) for all 2^24 triples, written to HOST memory */
OMPIMG_API int ompimg_selftest_luma_lut(uint8_t *lut_host, int mode);
/* out[m] = sat_u8(trunc(sqrt.approx(m + bias))) as the Sobel kernels compute it (bias 0.5 is the guard the
 * kernels use; it must equal min(255, isqrt(m)) for every m): HOST memory */
OMPIMG_API int ompimg_selftest_isqrt(uint8_t *out_host, unsigned int count, float bias);

/* ---- row-band decomposition of one large image over several devices of this process -------------
 * (BASELINE config 5).  The image is split into `ndev` bands of rows; each device holds its band
 * plus nothing else; Sobel reads its 1-row halos straight out of the neighbour device's band over
 * NVLink (peer loads inside the stencil kernel; no staging copy, no collective). */
typedef struct ompimg_bands ompimg_bands;
OMPIMG_API ompimg_bands *ompimg_bands_create(int width, int height, int channels, int ndev, const int *devices);
OMPIMG_API int  ompimg_bands_destroy(ompimg_bands *b);
OMPIMG_API int  ompimg_bands_rows(const ompimg_bands *b, int dev_index, int *row0, int *nrows);
OMPIMG_API int  ompimg_bands_synth(ompimg_bands *b, unsigned int seed);              /* fill every band on its device */
OMPIMG_API int  ompimg_bands_upload(ompimg_bands *b, const uint8_t *rgb_host);      /* dense width*height*channels */
OMPIMG_API int  ompimg_bands_rgb_to_edge(ompimg_bands *b, float *ms_max);           /* gray+Sobel on every band, halo via peer access */
OMPIMG_API int  ompimg_bands_download_edge(ompimg_bands *b, uint8_t *edge_host);    /* dense width*height plane, border bytes untouched */
OMPIMG_API int  ompimg_bands_checksum(ompimg_bands *b, unsigned long long *sum);    /* sum of all interior edge bytes */
/* a6 over bands (old/parallel_convolution.c:12-23): the banded 3-channel image is the pre-padded input, the output is
 * (height-K+1) x (width-K+1) x 3 uint8 with row r on the device that holds input row r.  The K-1 input rows below a band are
 * copied from the next device(s) over NVLink (cudaMemcpyPeerAsync) before the band's kernel starts: one exchange step. */
OMPIMG_API int  ompimg_bands_conv_u8(ompimg_bands *b, int K, const int32_t *taps, int32_t norm, float *ms_max);
OMPIMG_API int  ompimg_bands_download_conv(ompimg_bands *b, uint8_t *out_host);     /* dense (height-K+1) x (width-K+1) x 3 */
OMPIMG_API int  ompimg_bands_conv_checksum(ompimg_bands *b, unsigned long long *sum);   /* sum of all output bytes */
/* a7 over bands (old/parallel_avg_pixel.c:17-30): per-device partial channel sums, added on the host */
OMPIMG_API int  ompimg_bands_channel_sums(ompimg_bands *b, unsigned long long *sums3, float *ms_max);

/* one process per GPU (the launcher's form: torchrun, mpirun): the band above / below is another process's
 * allocation.  ompimg_ipc_export writes the 64-byte CUDA IPC handle of a buffer from ompimg_dev_alloc; the
 * neighbour passes it to ompimg_ipc_open and gets a device address valid in ITS process (peer access over
 * NVLink enabled lazily), to be used as row_above / row_below of ompimg_rgb_to_edge_band_dev or as the source
 * of ompimg_copy_d2d (a P2P copy).  The handle bytes travel by any side channel; no collective is involved. */
#define OMPIMG_IPC_HANDLE_BYTES 64
OMPIMG_API int   ompimg_ipc_export(const void *dev_ptr, unsigned char *handle_out /*[OMPIMG_IPC_HANDLE_BYTES]*/);
OMPIMG_API void *ompimg_ipc_open(const unsigned char *handle);     /* NULL on failure */
OMPIMG_API int   ompimg_ipc_close(void *mapped);

/* ---- JPEG decode in front of the kernels ---------------------------------------------------------
 * monolithic/src/main.c:21 decodes its input with the vendored stb_image v2.30 (stbi_load, req_comp 0) on one CPU core; on
 * `test.jpg` that takes longer than every kernel pass together.  Here the serial part (markers, Huffman, progressive passes)
 * runs on the host and yields quantised coefficients; dequantisation, IDCT, chroma upsampling and YCbCr->RGB run on the
 * GPU in stb_image's own integer arithmetic (stb_image.h:2430-2524, 3464-3527, 3659-3683), so the pixels equal the
 * reference's byte for byte.  Handles what stb's JPEG path handles for 8-bit Huffman files with 1 or 3 components and
 * sampling factors 1 and 2 (baseline and progressive, restart intervals); anything else is refused with OMPIMG_ERR_ARG.
 * Output: interleaved RGB (3 components) or the single plane, like stbi_load(..., 0). */
typedef struct ompimg_jpeg ompimg_jpeg;
OMPIMG_API int          ompimg_jpeg_probe(const uint8_t *bytes, size_t nbytes, int *width, int *height, int *channels);   /* header only */
OMPIMG_API ompimg_jpeg *ompimg_jpeg_parse(const uint8_t *bytes, size_t nbytes);      /* host only (no CUDA): entropy decoding; NULL on error */
OMPIMG_API void         ompimg_jpeg_free(ompimg_jpeg *j);
OMPIMG_API int          ompimg_jpeg_info(const ompimg_jpeg *j, int *width, int *height, int *channels);
OMPIMG_API int          ompimg_jpeg_to_rgb_dev(const ompimg_jpeg *j, uint8_t *out_dev, size_t pitch, int stream);     /* synchronises the stream */
OMPIMG_API int          ompimg_jpeg_to_rgb_host(const ompimg_jpeg *j, uint8_t *out_host);    /* dense width*height*channels */
/* main.c:21-35 as one call: decode, `passes` x convert_to_grayscale on the device, one download of width*height*channels bytes */
OMPIMG_API int          ompimg_jpeg_gray_host(const uint8_t *bytes, size_t nbytes, int passes, uint8_t *out_host, size_t out_capacity,
                                              int *width, int *height, int *channels);
/* for tests: component c's sampling factors, block grid, effective samples, coefficients [blocks_h][blocks_w][64] (natural
 * order, not dequantised) and its quantisation table; whether the three components are YCbCr (1) or already RGB (0) */
OMPIMG_API int          ompimg_jpeg_component(const ompimg_jpeg *j, int c, int *h_samp, int *v_samp, int *blocks_w, int *blocks_h,
                                              int *samples_x, int *samples_y, const int16_t **coeff, const uint16_t **dequant);
OMPIMG_API int          ompimg_jpeg_color_transform(const ompimg_jpeg *j);

/* ---- persistent worker mode (the reference's services exec the CLI once per request,
 * microservices/grayscale/app.py:24-34, event-driven/grayscale_service/app.py:54-65) -----------------
 * With OMPIMG_WORKER=<unix socket path> in the environment the two frozen prototypes do not create a
 * CUDA context (~3 s per fresh process): they hand the image to the process that runs
 * ompimg_worker_serve() (lib/ompimg_worker) and read the result back.  An unreachable worker means the
 * calling process initialises CUDA itself; there is still no CPU path.  Blocks until signalled. */
OMPIMG_API int ompimg_worker_serve(const char *socket_path, int device);

#ifdef __cplusplus
}
#endif
#endif /* OMPIMG_B200_H */
