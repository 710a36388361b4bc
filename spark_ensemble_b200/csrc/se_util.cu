// se_util.cu — fills, conversions and the counter-based synthetic generator (bench / tests).
#include "se_kernels.h"

namespace se {

namespace {

__global__ void __launch_bounds__(kBlock) fill_kernel(float* p, float v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    p[i] = v;
}

// Labels used as class indices (LogLoss, SAMME / SAMME.R) are validated ONCE per upload of the label slot, by this
// 4 B/row pass, instead of in every hot kernel (the per-row float->int->float round trip shares the XU pipe with
// ex2 / rcp: it cost the register LogLoss kernel 10-25 % and the SAMME.R tiles 5 %).  A label that is not an integer in
// [0, K) raises the flag; the reference throws for it (GBMLoss.scala:200-204, Classifier.validateLabel).
__global__ void __launch_bounds__(kBlock) validate_labels_kernel(const float* __restrict__ y, int64_t n, int K, int* bad) {
  bool b = false;
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    bool bb = false;
    (void)checked_label(y[i], K, bb);
    b = b || bb;
  }
  if (b && bad != nullptr) *reinterpret_cast<volatile int*>(bad) = 1;
}

// kind 0 uniform[a,b), 1 normal(mean a, sd b) (Box-Muller), 2 integer uniform in [a,b), 3 bernoulli(a)
__global__ void __launch_bounds__(kBlock) synth_kernel(float* p, int kind, uint64_t seed, float a, float b,
                                                      int64_t n, int64_t index_offset) {
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const uint64_t idx = (uint64_t)(i + index_offset);
    const uint64_t h1 = splitmix64(seed * 0x9E3779B97F4A7C15ull + 2 * idx);
    const float u1 = u01_from_bits(h1);
    float v;
    if (kind == 0) {
      v = a + (b - a) * u1;
    } else if (kind == 1) {
      const uint64_t h2 = splitmix64(seed * 0x9E3779B97F4A7C15ull + 2 * idx + 1);
      const float u2 = u01_from_bits(h2);
      const float r = sqrtf(-2.0f * logf(fmaxf(u1, 5.9604645e-8f)));
      v = a + b * r * cospif(2.0f * u2);
    } else if (kind == 2) {
      v = floorf(a + (b - a) * u1);
      if (v >= b) v = b - 1.0f;
    } else {
      v = (u1 < a) ? 1.0f : 0.0f;
    }
    p[i] = v;
  }
}

__global__ void __launch_bounds__(kBlock) f64_to_f32_kernel(const double* s, float* d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    d[i] = (float)s[i];
}

__global__ void __launch_bounds__(kBlock) scale_copy_kernel(const float* s, float* d, float scale, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    d[i] = s[i] * scale;
}

// order-preserving map float -> uint32 (negative values reversed, positives offset)
__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kBlock) radix_hist_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           int64_t n, uint32_t prefix, uint32_t mask, int shift,
                                                           double* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;  // kBlock == 256
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    float v = ld_stream1(a + i);
    if (b) v = fabsf(v - ld_stream1(b + i));
    const uint32_t key = float_key(v);
    if ((key & mask) == prefix) atomicAdd(&sh[(key >> shift) & 0xFFu], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (double)sh[threadIdx.x]);
}

// 32x32 shared-memory tile transpose: reads coalesced along the feature axis of the row-major chunk,
// writes coalesced along the row axis of the column-major matrix
__global__ void __launch_bounds__(256) transpose_rows_kernel(const float* __restrict__ src, int64_t rows, int d,
                                                             float* __restrict__ X, int64_t ld, int64_t row0) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j;
    const int c = c0 + tx;
    tile[j][tx] = (r < rows && c < d) ? src[r * d + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t r = r0 + tx;
    if (c < d && r < rows) X[(int64_t)c * ld + row0 + r] = tile[tx][j];
  }
}

inline int grid1(int64_t n, int sms) {
  int64_t need = (n + kBlock - 1) / kBlock;
  if (need < 1) need = 1;
  const int64_t cap = (int64_t)sms * 8;
  return (int)(need < cap ? need : cap);
}

}  // namespace

cudaError_t launch_radix_hist(const float* a, const float* b, int64_t n, uint32_t prefix, uint32_t mask,
                              int shift, double* hist, int sms, cudaStream_t s) {
  radix_hist_kernel<<<grid1(n, sms), kBlock, 0, s>>>(a, b, n, prefix, mask, shift, hist);
  return cudaGetLastError();
}
cudaError_t launch_transpose_rows(const float* src, int64_t rows, int d, float* X, int64_t ld, int64_t row0,
                                  cudaStream_t s) {
  if (rows <= 0 || d <= 0) return cudaSuccess;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((d + 31) / 32));
  transpose_rows_kernel<<<grid, 256, 0, s>>>(src, rows, d, X, ld, row0);
  return cudaGetLastError();
}
cudaError_t launch_validate_labels(const float* y, int64_t n, int K, int* bad, int sms, cudaStream_t s) {
  validate_labels_kernel<<<grid1(n, sms), kBlock, 0, s>>>(y, n, K, bad);
  return cudaGetLastError();
}
cudaError_t launch_fill(float* p, float v, int64_t n, int sms, cudaStream_t s) {
  fill_kernel<<<grid1(n, sms), kBlock, 0, s>>>(p, v, n);
  return cudaGetLastError();
}
cudaError_t launch_fill_synthetic(float* p, int kind, uint64_t seed, double a, double b, int64_t n,
                                  int64_t index_offset, int sms, cudaStream_t s) {
  synth_kernel<<<grid1(n, sms), kBlock, 0, s>>>(p, kind, seed, (float)a, (float)b, n, index_offset);
  return cudaGetLastError();
}
cudaError_t launch_f64_to_f32(const double* src, float* dst, int64_t n, int sms, cudaStream_t s) {
  f64_to_f32_kernel<<<grid1(n, sms), kBlock, 0, s>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t launch_scale_copy(const float* src, float* dst, float scale, int64_t n, int sms,
                              cudaStream_t s) {
  scale_copy_kernel<<<grid1(n, sms), kBlock, 0, s>>>(src, dst, scale, n);
  return cudaGetLastError();
}

}  // namespace se
