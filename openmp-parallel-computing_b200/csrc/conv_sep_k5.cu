// conv_sep_k5.cu -- the 5x5 instantiations of conv_sep.cuh (one translation unit per K: they compile in parallel)
#define OMPIMG_CONV_SEP_IMPL
#include "conv_sep.cuh"

namespace ompimg {
template cudaError_t launch_conv_sep_k<5>(const ConvSepParams &, int, int, int, dim3, int, cudaStream_t);
}
