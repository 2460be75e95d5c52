// conv_sep_k3.cu -- the 3x3 instantiations of conv_sep.cuh (one translation unit per K: they compile in parallel)
#define OMPIMG_CONV_SEP_IMPL
#include "conv_sep.cuh"

namespace ompimg {
template cudaError_t launch_conv_sep_k<3>(const ConvSepParams &, int, int, int, dim3, int, cudaStream_t);
}
