// tma.h -- the tensor-map encoder, shared by the kernels that are fed by cp.async.bulk.tensor copies (stencil.cu, conv_tma.cu).
#pragma once
#include <cuda.h>      // CUtensorMap and its enums (types only: the library links the static runtime, not libcuda)

namespace ompimg {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled, fetched from the driver through cudaGetDriverEntryPoint (stencil.cu); nullptr when the driver has none
EncodeTiledFn encode_tiled_fn();

}  // namespace ompimg
