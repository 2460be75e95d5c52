// kernels.h -- host-side launch interface between runtime.cu (the C ABI) and the kernel files.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include <atomic>

namespace ompimg {

extern std::atomic<long long> g_launch_count;   // every kernel launch of the library (runtime.cu)
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// pointwise.cu ------------------------------------------------------------------------------------
struct StreamParams {
    const uint8_t *in;
    uint8_t *out;
    size_t in_line_pitch, out_line_pitch;   // bytes between consecutive lines
    long long n_lines;
    uint32_t groups_per_line;               // full 16-pixel groups per line
    uint32_t chunks_per_line;               // ceil(groups/32): one chunk = one warp step
    uint32_t total_chunks;
    uint32_t step_lines, step_chunks;       // (grid warps) divmod chunks_per_line
    int tail_px;                            // pixels after the last full group of a line
    int *min_max;                           // a8 only
};

cudaError_t launch_gray_inplace(uint8_t *img, size_t pitch, int w, int h, int ch, int nframes, size_t fstride,
                                int fma_mode, int sm_count, cudaStream_t st);
cudaError_t launch_rgb_to_luma(const uint8_t *rgb, size_t rgb_pitch, int ch, uint8_t *luma, size_t luma_pitch, int w, int h,
                               int nframes, size_t rgb_stride, size_t luma_stride, int fma_mode, int sm_count, cudaStream_t st);
cudaError_t launch_replicate(const uint8_t *plane, size_t ppitch, uint8_t *img, size_t ipitch, int ch, int w, int h, int nframes,
                             size_t pstride, size_t istride, int sm_count, cudaStream_t st);
cudaError_t launch_gray_avg3(const uint8_t *rgb, size_t in_pitch, uint8_t *out, size_t out_pitch, int rows, int cols,
                             int *min_max_dev, int sm_count, cudaStream_t st);
cudaError_t launch_channel_sums(const uint8_t *rgb, size_t pitch, int rows, int cols, unsigned long long *sums_dev,
                                int sm_count, cudaStream_t st);
cudaError_t launch_synth(uint8_t *dst, unsigned long long nbytes, unsigned int seed, unsigned long long off,
                         int sm_count, cudaStream_t st);
cudaError_t launch_luma_lut(uint8_t *lut_dev, int fma_mode, cudaStream_t st);

// stencil.cu --------------------------------------------------------------------------------------
// One image (or a batch of equally shaped frames) through the 3x3 Sobel, either from a luma plane
// (channels == 1) or fused from interleaved RGB(A) (channels == 3 / 4).
struct StencilArgs {
    const uint8_t *src;      // plane or RGB(A)
    uint8_t *dst;            // edge plane
    size_t src_pitch, dst_pitch, src_stride, dst_stride;
    int width, height, channels, nframes;
    int fma_mode;
    // row-band mode (ompimg_bands_*): this launch covers global rows [row0, row0+height) of an image
    // `full_height` rows tall; the rows just above / below live in other allocations (peer devices).
    const uint8_t *halo_above;   // address of source row (row0-1), or nullptr
    const uint8_t *halo_below;   // address of source row (row0+height), or nullptr
    int row0, full_height;
    int no_bias;                 // 1: MUFU.SQRT was validated exact on this device, skip the +0.5 guard (runtime.cu)
};
cudaError_t launch_sobel(const StencilArgs &a, int sm_count, cudaStream_t st);
// a5 in place on 3-channel images (main_with_sobel.c:51-74 as a preparation kernel + one fused kernel that writes RGB):
// scratch is device memory of rows_bytes + border_bytes + entries_bytes, private to the launch until it has completed
struct PassScratch { size_t row_pitch, rows_bytes, border_bytes, entry_pitch, entries_bytes; size_t total() const { return rows_bytes + border_bytes + entries_bytes; } };
bool pass_inplace_supported(const uint8_t *img, size_t pitch, size_t stride, int w, int h, int ch, int nframes);
PassScratch pass_inplace_scratch(int w, int h, int nframes, int sm_count);
// `passes` iterations; scratch: sz.total() bytes for one pass, 2 * sz.total() for more
cudaError_t launch_gray_sobel_pass_inplace(uint8_t *img, size_t pitch, size_t stride, int w, int h, int nframes, const uint8_t *edge, size_t epitch,
                                           size_t estride, int fma_mode, int no_bias, int passes, uint8_t *scratch, int sm_count, cudaStream_t st);
int sqrt_exact_flag();          // runtime.cu: 1 when the bound device passed the exhaustive MUFU.SQRT check
cudaError_t launch_isqrt_selftest(uint8_t *out_dev, unsigned int count, float bias, cudaStream_t st);

// bands.cu: sum of the bytes of a width x rows region, added onto *sum_dev
cudaError_t launch_byte_sum(const uint8_t *src, size_t pitch, int width, int rows, unsigned long long *sum_dev, int sm_count, cudaStream_t st);

// conv.cu -----------------------------------------------------------------------------------------
cudaError_t launch_conv(const uint8_t *in, size_t in_pitch, int rows, int cols, int K, const int32_t *taps_host, int32_t norm,
                        void *out, size_t out_pitch, int out_is_u8, int sm_count, cudaStream_t st);

cudaError_t launch_accum_div(int32_t *out, size_t out_pitch, const int32_t *prior, size_t prior_pitch, int rows, int n, int32_t norm,
                             int sm_count, cudaStream_t st);

// conv_sep.cu: rank-1 taps u (vertical) x v (horizontal); false = outside its ranges, nothing launched
bool launch_conv_sep(const uint8_t *in, size_t in_pitch, int rows, int cols, int K, const int32_t *u, const int32_t *v, int32_t norm,
                     void *out, size_t out_pitch, int out_is_u8, int sm_count, cudaStream_t st, cudaError_t *err);

}  // namespace ompimg
