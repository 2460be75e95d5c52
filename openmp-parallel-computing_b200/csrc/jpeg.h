// jpeg.h -- interface between the C ABI (runtime.cu) and the JPEG decode step (jpeg.cu)
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace ompimg {

struct JpegComp {
    int id = 0, h = 1, v = 1, tq = 0;
    int x = 0, y = 0;                 // effective samples of the component
    int bw = 0, bh = 0;               // 8x8 blocks per row / column, MCU padding included
    std::vector<int16_t> coeff;       // [bh][bw][64] quantised coefficients in natural order (not yet dequantised)
};
struct JpegCoeffs {
    int width = 0, height = 0, ncomp = 0;
    bool progressive = false;
    int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    bool jfif = false, ids_are_rgb = false;
    int app14_transform = -1;
    uint16_t dequant[4][64];          // natural order
    JpegComp comp[3];
    int out_channels() const { return ncomp == 3 ? 3 : 1; }
};

// host: everything serial (markers, Huffman, progressive passes) -> coefficients
bool jpeg_entropy_decode(const uint8_t *data, size_t size, JpegCoeffs &out, std::string &err);
// host: dimensions only (SOF); false if the bytes are not a JPEG this decoder handles
bool jpeg_probe(const uint8_t *data, size_t size, int *w, int *h, int *channels);
// device: dequantise + IDCT + upsample + colour conversion into `out` (RGB interleaved or a single plane)
size_t jpeg_device_scratch_bytes(const JpegCoeffs &jc);
cudaError_t launch_jpeg_backend(const JpegCoeffs &jc, uint8_t *scratch, uint8_t *out, size_t out_pitch, int sm_count, cudaStream_t st);

}  // namespace ompimg
