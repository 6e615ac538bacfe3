// TMA row gather for sm_100a: tensor-map construction (driver encoder fetched through the runtime, so the library
// does not link libcuda) and the `tile::gather4` instruction.
//
// Measured on B200 (tools/tma_gather_probe.cu): the tensor-map box must be ONE row (the instruction names 4 row
// coordinates; a 4-row box raises an illegal-instruction fault); a warp issues one gather4 per ~46 ns regardless of
// how many are in flight, issuing warps scale linearly until the SM saturates at ~6.5 ns per 512-byte gather4
// (11.6 TB/s over 148 SMs); single-instruction latency 0.44 us; rows that are out of range are zero-filled at ~9 ns
// each (slower than fetching), and many lanes gathering the SAME row hot-spot one L2 slice (0.6 TB/s).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encoder() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// 2-D row-major tensor [n_rows, n_cols] with row pitch `pitch_bytes`; box = box_cols x 1 row (the gather4 box)
inline bool make_row_gather_map(CUtensorMap* tm, const void* base, int64_t n_rows, int64_t n_cols, int64_t pitch_bytes,
                                CUtensorMapDataType dtype, int box_cols, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = encoder();
  if (!enc || n_rows < 1) return false;
  cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 1};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// rows r0..r3 (any order, any distance) x [col, col + box_cols) → 4 consecutive box rows at dst_smem; completion as
// complete_tx bytes on `bar`
__device__ __forceinline__ void gather4(uint32_t dst_smem, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                        uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst_smem), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
#endif

}  // namespace tma
