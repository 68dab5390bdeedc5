// CUtensorMap construction for the kernels that use tiled TMA copies (2-D tensor loads / stores).
// cuTensorMapEncodeTiled is reached through the runtime's driver entry point, so the library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace gatres {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}
// row-major fp32 [rows][cols] in boxes of box_rows x box_cols; swizzle128: the box's shared-memory image is SWIZZLE_128B
// (box_cols * 4 must be 128 then), else dense rows
static inline bool make_map_2d(CUtensorMap* map, const float* ptr, unsigned long long rows, unsigned cols, unsigned box_cols,
                               unsigned box_rows, bool swizzle128) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[2] = {cols, rows}, strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gatres
