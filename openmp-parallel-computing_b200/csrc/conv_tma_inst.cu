// conv_tma_inst.cu -- explicit instantiation of one group of conv_tma.cuh's kernel variants: compiled once per
// (CT_INST_K in 3 5 7 9) x (CT_INST_PART in 0..3) by the Makefile, so that the sixteen groups build in parallel.
#ifndef CT_INST_K
#define CT_INST_K 3      // a bare `nvcc -c` of this file builds the smallest group
#define CT_INST_PART 0
#endif
#define OMPIMG_CONV_TMA_IMPL
#include <cstdlib>

#include "conv_tma.cuh"

namespace ompimg {
template cudaError_t launch_conv_tma_part<CT_INST_K, CT_INST_PART>(ConvTmaParams &, int, int, int, cudaStream_t);
}
