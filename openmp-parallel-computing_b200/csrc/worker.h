// worker.h -- client side of the persistent worker mode (worker.cu), used by the frozen prototypes in runtime.cu.
#pragma once
#include <cstddef>

namespace ompimg {
// 1: served by the worker named by OMPIMG_WORKER, 0: no worker configured / reachable (caller runs in-process),
// < 0: the worker answered with an error (message in err)
int worker_call_gray(unsigned char *data, int w, int h, int ch, int luma_mode, char *err, size_t err_cap);
int worker_call_sobel(const unsigned char *src, unsigned char *dst, int w, int h, char *err, size_t err_cap);
int worker_call_jpeg_gray(const unsigned char *jpeg, size_t nbytes, int passes, int luma_mode, unsigned char *out, size_t out_bytes, int w, int h, int ch,
                          char *err, size_t err_cap);
}  // namespace ompimg
