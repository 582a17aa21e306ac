// tmap.cuh — host-side construction of 2-D TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point:
// the library links no libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gssd {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// matrix [rows, cols] (cols innermost) of `esize`-byte elements, box = box_rows x box_cols, zero fill outside the tensor
inline int make_tmap_2d(CUtensorMap *map, CUtensorMapDataType dt, int esize, const void *base, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return (int)cudaErrorNotSupported;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * (uint64_t)esize};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace gssd
