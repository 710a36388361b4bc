// se_tma.cuh — TMA (cp.async.bulk.tensor) + mbarrier building blocks shared by the tiled kernels: a [K][ld] fp32
// class-major array is described by a 2-D tensor map whose boxes are [K][tile_rows]; one instruction brings a whole
// tile into shared memory, rows past n are zero-filled by the TMA unit, completion is counted on an mbarrier.
#pragma once

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

namespace se {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one box [K][tile_rows] of the [K][ld] array, first row `row0`, into shared memory [K][tile_rows]
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, int row0, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(row0), "r"(0), "r"(smem_u32(bar))
      : "memory");
}
// box starting at (row0, first_row) of a [rows][ld] array
__device__ __forceinline__ void tma_load_tile_at(void* dst, const CUtensorMap* map, int row0, int first_row,
                                                 uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(row0), "r"(first_row), "r"(smem_u32(bar))
      : "memory");
}
// generic-proxy accesses to a stage are ordered before the async-proxy (TMA) refill of the same bytes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// ---- host side --------------------------------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// general form: the array has `rows` rows of stride ld; boxes are [box_rows][tile_cols]
inline cudaError_t make_tile_map_rows(CUtensorMap* m, const float* base, int64_t n, int64_t ld, int64_t rows,
                                      int tile_cols, int box_rows) {
  auto enc = tensor_map_encoder();
  if (enc == nullptr) return cudaErrorNotSupported;
  const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)rows};
  // a one-row array has no meaningful stride (1-D slots are unpadded): any multiple of 16 B satisfies the encoder
  const cuuint64_t gstride[1] = {rows > 1 ? (cuuint64_t)ld * sizeof(float) : (((cuuint64_t)ld * sizeof(float) + 15) / 16) * 16};
  const cuuint32_t box[2] = {(cuuint32_t)tile_cols, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
// [K][ld] fp32, rows [0, n) valid: boxes of [K][tile_rows]; reads past n are zero-filled
inline cudaError_t make_tile_map(CUtensorMap* m, const float* base, int64_t n, int64_t ld, int K, int tile_rows) {
  return make_tile_map_rows(m, base, n, ld, K, tile_rows, K);
}

}  // namespace se
