// se_gbm_staged.cu — LogLoss(K) GBM kernels for wide K: shared-memory staging with TMA bulk copies.
//
// Reference arithmetic: boosting/GBMLoss.scala:196-263 (LogLoss), :50-74 (aggregator),
// classification/GBMClassifier.scala:337-375 (residuals), :437-449 (update).
//
// The register-resident kernel of se_gbm.cu keeps all K classes of a row in registers; beyond K = 8 that
// needs 255 registers, spills and runs one CTA per SM (ncu: 12 % occupancy, ~1100 instructions per row,
// 0.3 of the HBM roofline at K = 26).  Here a CTA stages a tile of R rows x K classes of F and h in shared
// memory with `cp.async.bulk` (TMA, one 4R-byte copy per class row, completion on an mbarrier), double
// buffered: the next tile's 2K bulk copies are in flight while the current tile is evaluated.  A thread owns
// one row of the tile and walks the classes three times out of shared memory (max / sum-exp / outputs), so
// only the per-class fp64 accumulators live in registers.  Layout [K][ld] fp32 is unchanged.
#include <stdlib.h>

#include "se_kernels.h"
#include "se_loss.cuh"

namespace se {

namespace {

constexpr int kRows = 128;   // rows per tile == threads per CTA
constexpr int kStages = 2;

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
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE>
struct StagedTraits {
  static constexpr bool kReadH = (MODE == GBM_EVAL || MODE == GBM_UPDATE || MODE == GBM_UPDATE_RESID ||
                                  MODE == GBM_UPDATE_NEWTON);
  static constexpr bool kWriteF = (MODE == GBM_UPDATE || MODE == GBM_UPDATE_RESID || MODE == GBM_UPDATE_NEWTON);
  static constexpr bool kNewton = (MODE == GBM_RESID_NEWTON || MODE == GBM_UPDATE_NEWTON);
  static constexpr bool kWriteR = (MODE == GBM_RESID || MODE == GBM_UPDATE_RESID || kNewton);
  static constexpr bool kSumLoss = (MODE == GBM_EVAL || kWriteF || MODE == GBM_MEAN_LOSS);
  static constexpr bool kPerClassAcc = (MODE == GBM_EVAL || kNewton);
  static constexpr bool kReduce = kSumLoss || kNewton;
};

// dynamic shared memory: [stage][ {F,h} ][K][kRows] floats, then the mbarriers
template <int KMAX, int MODE>
__global__ void __launch_bounds__(kRows) gbm_logloss_staged_kernel(const GbmArgs a) {
  using T = StagedTraits<MODE>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = a.dim;
  constexpr int kArrays = T::kReadH ? 2 : 1;
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  const int stage_floats = kArrays * K * kRows;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)(a.stages >= 2 ? 2 : 1) * stage_floats);
  __shared__ float s_coef[kMaxDim];

  const int tid = threadIdx.x;
  if (tid < K) s_coef[tid] = a.coef[tid];
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t ntiles = (a.n + kRows - 1) / kRows;
  const int64_t ld = a.ld;
  const bool has_w = (a.w != nullptr);

  auto issue = [&](int64_t tile, int stage) {  // one elected thread
    const int64_t row0 = tile * kRows;
    int64_t rr = a.n - row0;
    if (rr > kRows) rr = kRows;
    const uint32_t bytes = (uint32_t)(((rr + 3) & ~int64_t(3)) * sizeof(float));  // rows padded to 16 B (ld is padded)
    float* dst = stage_base + (size_t)stage * stage_floats;
    mbar_expect_tx(&bars[stage], bytes * (uint32_t)(kArrays * K));
    for (int k = 0; k < K; ++k) {
      tma_load_1d(dst + k * kRows, a.F + k * ld + row0, bytes, &bars[stage]);
      if (T::kReadH) tma_load_1d(dst + (K + k) * kRows, a.h + k * ld + row0, bytes, &bars[stage]);
    }
  };

  constexpr int NRED = T::kPerClassAcc ? KMAX + 1 : 1;
  double acc_loss = 0.0;
  // per-class sums (Σ h_k g_k or Σ max(H_k,1e-2)): each thread drops its row's contribution into its own
  // shared-memory slot (the F slot it has just consumed), and the 4 warps then sum the tile's 128 rows per
  // class with shuffles into fp64 accumulators in shared memory — no per-thread per-class registers.
  __shared__ double s_acc[KMAX];
  if (tid < KMAX) s_acc[tid] = 0.0;

  const bool two_stage = (a.stages >= 2);
  int64_t tile = blockIdx.x;
  if (two_stage && tid == 0 && tile < ntiles) issue(tile, 0);
  uint32_t it = 0;
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = two_stage ? (it & 1) : 0;
    const int64_t next = tile + gridDim.x;
    if (two_stage) {
      // the other stage was fully consumed before the __syncthreads() that ended the previous iteration
      if (tid == 0 && next < ntiles) issue(next, stage ^ 1);
    } else if (tid == 0) {
      issue(tile, 0);  // single stage: other resident CTAs of the SM cover this tile's load latency
    }
    const int64_t row = tile * kRows + tid;
    const bool in = row < a.n;
    const float yv = in ? ld_stream1(a.y + row) : 0.f;
    const float wv = (T::kNewton && has_w && in) ? ld_stream1(a.w + row) : 1.0f;
    const float cv = (T::kPerClassAcc && a.bag != nullptr && in) ? ld_stream1(a.bag + row) : 1.0f;  // bag multiplicity
    mbar_wait(&bars[stage], two_stage ? ((it >> 1) & 1) : (it & 1));
    float* sF = stage_base + (size_t)stage * stage_floats + tid;
    const float* sH = sF + K * kRows;
    const int yi = (int)yv;

    // pass A (online soft-max, one sweep over the classes): running max m and s = Σ_k exp(p_k - m), rescaled
    // whenever the max moves — no separate max pass; p_y picked up on the way.
    float m = -INFINITY, ssum = 0.f, py = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float p = T::kReadH ? fmaf(s_coef[k], sH[k * kRows], sF[k * kRows]) : sF[k * kRows];  // GBMLoss.scala:56-59
      const float mn = fmaxf(m, p);
      ssum = fmaf(ssum, ex2_approx((m - mn) * kLog2e), ex2_approx((p - mn) * kLog2e));
      m = mn;
      py = (k == yi) ? p : py;
    }
    // ssum >= 1 (the max term contributes exactly 1): log Σ exp(p_k) = m + log1p(ssum - 1)
    const float lse = m + log1p_pos(ssum - 1.0f);
    const float inv_s = rcp_approx(ssum);
    if (T::kSumLoss && in) acc_loss += (double)(((MODE == GBM_EVAL) ? cv : 1.0f) * (lse - py));  // GBMLoss.scala:206-221
    // pass B: per-class outputs (no per-class registers: sums go through shared memory)
    if (T::kPerClassAcc || T::kWriteF || T::kWriteR) {
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        {
          const float hk = T::kReadH ? sH[k * kRows] : 0.f;
          const float p = T::kReadH ? fmaf(s_coef[k], hk, sF[k * kRows]) : sF[k * kRows];
          const float sm = ex2_approx((p - m) * kLog2e) * inv_s;   // exp(p_k - lse)
          const float gk = sm - ((k == yi) ? 1.0f : 0.0f);         // :223-238
          if (MODE == GBM_EVAL) sF[k * kRows] = in ? cv * hk * gk : 0.f;  // :66-72 (summed per class below)
          if (T::kWriteF && in) a.F[k * ld + row] = p;
          if (T::kNewton) {
            const float hc = fmaxf(sm * (1.0f - sm), 1e-2f);       // :240-256, GBMClassifier.scala:342
            sF[k * kRows] = in ? cv * hc : 0.f;
            if (in) {
              a.r[k * ld + row] = -gk / hc;                        // :362
              a.wout[k * ld + row] = 0.5f * hc * wv;               // :364 (x 1/S_k later)
            }
          } else if (T::kWriteR) {
            if (in) a.r[k * ld + row] = -gk;                       // :371
          }
        }
      }
    }
    if (T::kPerClassAcc) {
      __syncthreads();  // all 128 contributions of every class are in shared memory
      const int lane = tid & 31, warp = tid >> 5;
      const float* c0 = stage_base + (size_t)stage * stage_floats;
      for (int k = warp; k < K; k += kRows / 32) {
        const float* c = c0 + k * kRows;
        float v = (c[lane] + c[lane + 32]) + (c[lane + 64] + c[lane + 96]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) s_acc[k] += (double)v;
      }
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }

  if (T::kReduce) {
    // Σloss: block reduction over 4 warps; per-class sums are already block-level in s_acc.
    __shared__ double sm_red[kRows / 32];
    __shared__ double s_tot[NRED];
    __shared__ bool is_last;
    const int lane = tid & 31, warp = tid >> 5;
    {
      const double v = warp_sum(acc_loss);
      if (lane == 0) sm_red[warp] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double v = 0.0;
      for (int w = 0; w < kRows / 32; ++w) v += sm_red[w];
      a.ws.partials[(size_t)blockIdx.x * NRED] = v;
    }
    if (T::kPerClassAcc && tid < KMAX) a.ws.partials[(size_t)blockIdx.x * NRED + 1 + tid] = (tid < K) ? s_acc[tid] : 0.0;
    __threadfence();  // every writer publishes its partials before the ticket is taken
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const unsigned int ticket = atomicInc(a.ws.counter, gridDim.x - 1);
      is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed-order cross-CTA reduction: warp w handles outputs w, w+4, ...; lanes stride over the CTAs
    for (int k = warp; k < NRED; k += kRows / 32) {
      double v = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) v += __ldcg(&a.ws.partials[(size_t)b * NRED + k]);
      v = warp_sum(v);
      if (lane == 0) s_tot[k] = v;  // [0] = Σloss, [1 + k] per class (register-kernel convention)
    }
    __syncthreads();
    peer_exchange(s_tot, NRED, a.ws);
  }
}

template <int KMAX>
cudaError_t launch_staged_k(int mode, const GbmArgs& a, int sms, cudaStream_t st) {
  const int K = a.dim;
  const bool read_h = (mode == GBM_EVAL || mode == GBM_UPDATE || mode == GBM_UPDATE_RESID || mode == GBM_UPDATE_NEWTON);
  const size_t stage_bytes = (size_t)(read_h ? 2 : 1) * K * kRows * sizeof(float);
  static const int forced_stages = [] { const char* e = getenv("SE_LOGLOSS_STAGES"); return e ? atoi(e) : 0; }();
  // wide rows: one stage and twice the resident CTAs beats two stages (occupancy is shared-memory bound)
  const int stages = forced_stages ? (forced_stages >= 2 ? 2 : 1) : (stage_bytes > 16 * 1024 ? 1 : 2);
  GbmArgs args = a;
  args.stages = stages;
  const size_t smem = stages * stage_bytes + kStages * sizeof(uint64_t) + 16;
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm < 1) return cudaErrorInvalidValue;
  if (per_sm > 8) per_sm = 8;
  int64_t need = (a.n + kRows - 1) / kRows;
  if (need < 1) need = 1;
  int64_t cap = (int64_t)per_sm * sms;
  if (cap > kMaxGridPartials) cap = kMaxGridPartials;
  const int grid = (int)(need < cap ? need : cap);
#define SE_CASE(M)                                                                                   \
  case M: {                                                                                          \
    auto kern = gbm_logloss_staged_kernel<KMAX, M>;                                                  \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                  \
    kern<<<grid, kRows, smem, st>>>(args);                                                              \
    break;                                                                                           \
  }
  switch (mode) {
    SE_CASE(GBM_RESID)
    SE_CASE(GBM_RESID_NEWTON)
    SE_CASE(GBM_EVAL)
    SE_CASE(GBM_UPDATE)
    SE_CASE(GBM_UPDATE_RESID)
    SE_CASE(GBM_UPDATE_NEWTON)
    SE_CASE(GBM_MEAN_LOSS)
    default: return cudaErrorInvalidValue;
  }
#undef SE_CASE
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gbm_logloss_staged(int mode, const GbmArgs& a, int sms, cudaStream_t st) {
  const int K = a.dim;
  if (K < 1 || K > kMaxDim) return cudaErrorInvalidValue;
  if (K <= 8) return launch_staged_k<8>(mode, a, sms, st);
  if (K <= 16) return launch_staged_k<16>(mode, a, sms, st);
  return launch_staged_k<32>(mode, a, sms, st);
}

}  // namespace se
