// conv_sep_inst.cu -- explicit instantiation of one group of conv_sep.cuh's kernel variants: compiled once per
// (CS_INST_K in 3 5 7 9) x (CS_INST_PART in 0..7) by the Makefile, so that the groups build in parallel.
#ifndef CS_INST_K
#define CS_INST_K 3      // a bare `nvcc -c` of this file builds the smallest group
#define CS_INST_PART 0
#endif
#define OMPIMG_CONV_SEP_IMPL
#include "conv_sep.cuh"

namespace ompimg {
template cudaError_t launch_conv_sep_part<CS_INST_K, CS_INST_PART>(const ConvSepParams &, int, dim3, int, cudaStream_t);
}
