// conv_tma_k3.cu -- the 3x3 instantiations of conv_tma.cuh (one translation unit per K: they compile in parallel)
#define OMPIMG_CONV_TMA_IMPL
#include <cstdlib>

#include "conv_tma.cuh"

namespace ompimg {
template cudaError_t launch_conv_tma_k<3>(ConvTmaParams &, int, int, int, int, int, cudaStream_t);
}
