// conv_tma_k5.cu -- the 5x5 instantiations of conv_tma.cuh (one translation unit per K: they compile in parallel)
#define OMPIMG_CONV_TMA_IMPL
#include <cstdlib>

#include "conv_tma.cuh"

namespace ompimg {
template cudaError_t launch_conv_tma_k<5>(ConvTmaParams &, int, int, int, int, int, cudaStream_t);
}
