// se_models.cu — on-device base-model evaluation over the column-major feature matrix X[d][n].
//
// The reference obtains the direction h (GBMRegressor.scala:405,435), class probabilities and
// predictions (BoostingClassifier.scala:199-200,233) by calling third-party Spark ML
// `model.predict(features)` once per row on the JVM.  Keeping X resident in HBM and evaluating the
// fitted model here means h never crosses PCIe.  Supported: binary decision trees
// with continuous splits (Spark ContinuousSplit.shouldGoLeft: x <= threshold goes left) and linear
// models.  HasSubBag.slice (ensemble/HasSubBag.scala:81-84) is folded into the feature->column map.
#include "se_kernels.h"

namespace se {

namespace {

constexpr int TV = 4;  // rows per thread (one float4 of outputs)

// Tree arrays are staged once per CTA in shared memory; each thread walks TV rows in lockstep so
// TV independent gathers are in flight.  Rows of a warp are consecutive, so every X access of a
// level is a coalesced 128 B segment per distinct feature.
__global__ void __launch_bounds__(kBlock) tree_predict_kernel(const TreeArgs a) {
  extern __shared__ unsigned char smem_raw[];
  int32_t* s_feat = reinterpret_cast<int32_t*>(smem_raw);
  float* s_thr = reinterpret_cast<float*>(s_feat + a.n_nodes);
  int32_t* s_left = reinterpret_cast<int32_t*>(s_thr + a.n_nodes);
  int32_t* s_right = s_left + a.n_nodes;
  float* s_val = reinterpret_cast<float*>(s_right + a.n_nodes);
  const bool scalar = (a.n_out == 1);
  for (int i = threadIdx.x; i < a.n_nodes; i += kBlock) {
    s_feat[i] = a.feature[i];
    s_thr[i] = a.threshold[i];
    s_left[i] = a.left[i];
    s_right[i] = a.right[i];
    if (scalar) s_val[i] = a.value[i];
  }
  __syncthreads();
  const int64_t ngroups = (a.n + TV - 1) / TV;
  for (int64_t g = (int64_t)blockIdx.x * kBlock + threadIdx.x; g < ngroups;
       g += (int64_t)gridDim.x * kBlock) {
    const int64_t i0 = g * TV;
    int node[TV];
    bool live[TV];
    bool any = false;
#pragma unroll
    for (int e = 0; e < TV; ++e) {
      node[e] = 0;
      live[e] = (i0 + e < a.n) && (s_feat[0] >= 0);
      any |= live[e];
    }
    while (any) {
      float x[TV];
#pragma unroll
      for (int e = 0; e < TV; ++e)
        if (live[e]) x[e] = __ldg(a.X + (int64_t)s_feat[node[e]] * a.ld + i0 + e);
      any = false;
#pragma unroll
      for (int e = 0; e < TV; ++e)
        if (live[e]) {
          node[e] = (x[e] <= s_thr[node[e]]) ? s_left[node[e]] : s_right[node[e]];
          live[e] = s_feat[node[e]] >= 0;
          any |= live[e];
        }
    }
    if (!scalar) {
      // leaf vectors (class probabilities of a classification tree): one coalesced row store per class
      for (int k = 0; k < a.n_out; ++k) {
        if (i0 + TV <= a.n) {
          st_stream4(a.out + (int64_t)k * a.ld_out + i0,
                     make_float4(__ldg(a.value + node[0] * a.n_out + k), __ldg(a.value + node[1] * a.n_out + k),
                                 __ldg(a.value + node[2] * a.n_out + k), __ldg(a.value + node[3] * a.n_out + k)));
        } else {
#pragma unroll
          for (int e = 0; e < TV; ++e)
            if (i0 + e < a.n) a.out[(int64_t)k * a.ld_out + i0 + e] = __ldg(a.value + node[e] * a.n_out + k);
        }
      }
    } else if (i0 + TV <= a.n) {
      st_stream4(a.out + i0, make_float4(s_val[node[0]], s_val[node[1]], s_val[node[2]], s_val[node[3]]));
    } else {
#pragma unroll
      for (int e = 0; e < TV; ++e)
        if (i0 + e < a.n) a.out[i0 + e] = s_val[node[e]];
    }
  }
}

constexpr int LU = 8;

__global__ void __launch_bounds__(kBlock) linear_predict_kernel(const float* __restrict__ X, int64_t n,
                                                               int64_t ld, int n_coef,
                                                               const float* __restrict__ coef,
                                                               const int32_t* __restrict__ cols,
                                                               float intercept, float* __restrict__ out) {
  const int64_t n4 = n >> 2;
  for (int64_t g = (int64_t)blockIdx.x * kBlock + threadIdx.x; g < n4;
       g += (int64_t)gridDim.x * kBlock) {
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    for (int j0 = 0; j0 < n_coef; j0 += LU) {
      float4 v[LU];
      float c[LU];
#pragma unroll
      for (int u = 0; u < LU; ++u)
        if (j0 + u < n_coef) {
          const int64_t col = cols ? cols[j0 + u] : (j0 + u);
          v[u] = ld_stream4(X + col * ld + 4 * g);
          c[u] = coef[j0 + u];
        }
#pragma unroll
      for (int u = 0; u < LU; ++u)
        if (j0 + u < n_coef) {
          float4& s = (u & 1) ? s1 : s0;
          s.x = fmaf(c[u], v[u].x, s.x); s.y = fmaf(c[u], v[u].y, s.y);
          s.z = fmaf(c[u], v[u].z, s.z); s.w = fmaf(c[u], v[u].w, s.w);
        }
    }
    st_stream4(out + 4 * g, make_float4(intercept + (s0.x + s1.x), intercept + (s0.y + s1.y),
                                        intercept + (s0.z + s1.z), intercept + (s0.w + s1.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float s = 0.f;
    for (int j = 0; j < n_coef; ++j) {
      const int64_t col = cols ? cols[j] : j;
      s = fmaf(coef[j], X[col * ld + i], s);
    }
    out[i] = intercept + s;
  }
}

}  // namespace

cudaError_t launch_tree_predict(const TreeArgs& a, int sms, cudaStream_t st) {
  const size_t smem = (size_t)a.n_nodes * 5 * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(tree_predict_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int64_t ngroups = (a.n + TV - 1) / TV;
  int64_t need = (ngroups + kBlock - 1) / kBlock;
  if (need < 1) need = 1;
  const int64_t cap = (int64_t)sms * 8;
  tree_predict_kernel<<<(int)(need < cap ? need : cap), kBlock, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_linear_predict(const float* X, int64_t n, int64_t ld, int n_coef,
                                  const float* coef, const int32_t* cols, float intercept,
                                  float* out, int sms, cudaStream_t st) {
  int64_t need = ((n >> 2) + kBlock - 1) / kBlock;
  if (need < 1) need = 1;
  const int64_t cap = (int64_t)sms * 8;
  linear_predict_kernel<<<(int)(need < cap ? need : cap), kBlock, 0, st>>>(X, n, ld, n_coef, coef, cols,
                                                                        intercept, out);
  return cudaGetLastError();
}

}  // namespace se
