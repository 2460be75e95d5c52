// conv_sep_k9.cu -- the 9x9 instantiations of conv_sep.cuh (one translation unit per K: they compile in parallel)
#define OMPIMG_CONV_SEP_IMPL
#include "conv_sep.cuh"

namespace ompimg {
template cudaError_t launch_conv_sep_k<9>(const ConvSepParams &, int, int, int, dim3, int, cudaStream_t);
}
