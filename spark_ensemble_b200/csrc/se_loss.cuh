// se_loss.cuh — per-row fp32 loss / gradient / hessian evaluators for the GBM kernels.
//
// Device restatement of boosting/GBMLoss.scala:129-318 (reference is fp64 on the JVM).  Rows are
// stored fp32 in HBM and evaluated in fp32 with numerically stable forms so that sums (accumulated
// in fp64) and per-row outputs stay within 1e-5 relative of the reference wherever the reference
// itself is finite.  `y` is the RAW label; the bernoulli/exponential encoding 2y-1 (:272,:297) is
// applied here.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "../../include/se_abi.h"

namespace se {

struct LGH {
  float l, g, h;
};

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : x); }

// ---- SFU-based building blocks -------------------------------------------------------------------
// The bernoulli / logcosh / logloss / SAMME.R kernels are transcendental-heavy; with libm-accurate
// expf/log1pf/division they become issue-bound (~100 instr/row) well below the HBM roofline.  These
// forms use one MUFU op each plus a few FMAs and keep relative error <= ~3e-7 (inside the 1e-5 bar).
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// e^x with the rounding error of x*log2(e) compensated (matters when |x| is large and e^x is an output)
__device__ __forceinline__ float exp_fast(float x) {
  const float a = x * kLog2e;
  const float lo = fmaf(x, kLog2e, -a) + x * 1.925963033500011e-8f;  // product residue + log2e tail
  const float e = ex2_approx(a);
  return fmaf(e, lo * kLn2, e);
}
// e^x, x <= 0, as an intermediate (argument error |x|*6e-8 only scales an already tiny result)
__device__ __forceinline__ float exp_neg_fast(float x) { return ex2_approx(x * kLog2e); }

// log(1+t) for t in [0,1] to ~1e-7 RELATIVE accuracy (also for tiny t): 2*atanh(t/(2+t))
__device__ __forceinline__ float log1p_unit(float t) {
  const float s = t * rcp_approx(2.0f + t);  // in [0, 1/3]
  const float s2 = s * s;
  float p = 1.0f / 13.0f;
  p = fmaf(p, s2, 1.0f / 11.0f);
  p = fmaf(p, s2, 1.0f / 9.0f);
  p = fmaf(p, s2, 1.0f / 7.0f);
  p = fmaf(p, s2, 1.0f / 5.0f);
  p = fmaf(p, s2, 1.0f / 3.0f);
  p = fmaf(p, s2, 1.0f);
  return 2.0f * s * p;
}
// log(1+t) for any t >= 0
__device__ __forceinline__ float log1p_pos(float t) {
  return (t <= 1.0f) ? log1p_unit(t) : lg2_approx(1.0f + t) * kLn2;
}
// log(x) for x > 0 (relative error ~2.4e-7 away from x = 1, absolute 1.7e-7 near it)
__device__ __forceinline__ float log_fast(float x) { return lg2_approx(x) * kLn2; }

// tanh(d), log(cosh(d)) and 1/cosh^2(d) together: no overflow, no cancellation near 0 or at |d| >> 1
__device__ __forceinline__ void tanh_logcosh(float d, float& th, float& lc, float& sech2) {
  const float a = fabsf(d);
  const float x2 = d * d;
  // |d| < 0.25: Taylor series (next terms < 1e-7 relative)
  const float th_s = d * fmaf(x2, fmaf(x2, fmaf(x2, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
  const float lc_s = x2 * fmaf(x2, fmaf(x2, fmaf(x2, -17.0f / 2520.0f, 1.0f / 45.0f), -1.0f / 12.0f), 0.5f);
  // otherwise through t = e^{-2|d|}: tanh = (1-t)/(1+t), sech^2 = 4t/(1+t)^2
  const float t = exp_neg_fast(-2.0f * a);
  const float inv = rcp_approx(1.0f + t);
  const float th_e = copysignf((1.0f - t) * inv, d);
  const float lc_e = (a - kLn2) + log1p_unit(t);
  const bool small = a < 0.25f;
  th = small ? th_s : th_e;
  lc = small ? lc_s : lc_e;
  sech2 = small ? (1.0f - th_s * th_s) : 4.0f * t * inv * inv;
}

template <int LOSS>
__device__ __forceinline__ LGH eval_loss(float y, float p, float param) {
  LGH o;
  if constexpr (LOSS == SE_LOSS_SQUARED) {  // :129-137
    const float d = y - p;
    o.l = 0.5f * d * d;
    o.g = -d;
    o.h = 1.0f;
  } else if constexpr (LOSS == SE_LOSS_ABSOLUTE) {  // :139-143
    const float d = y - p;
    o.l = fabsf(d);
    o.g = -sgnf(d);
    o.h = 0.f;
  } else if constexpr (LOSS == SE_LOSS_HUBER) {  // :168-177
    const float d = y - p, a = fabsf(d);
    const bool in = a <= param;
    o.l = in ? 0.5f * d * d : param * (a - 0.5f * param);
    o.g = in ? -d : -param * sgnf(d);
    o.h = 0.f;
  } else if constexpr (LOSS == SE_LOSS_QUANTILE) {  // :179-188
    const float d = y - p;
    const bool up = y > p;
    o.l = up ? param * d : (param - 1.0f) * d;
    o.g = up ? -param : (1.0f - param);
    o.h = 0.f;
  } else if constexpr (LOSS == SE_LOSS_LOGCOSH || LOSS == SE_LOSS_SCALED_LOGCOSH) {  // :145-166
    const float d = y - p;
    float t, lc, sech2;
    tanh_logcosh(d, t, lc, sech2);
    float s = 1.0f;
    if constexpr (LOSS == SE_LOSS_SCALED_LOGCOSH) s = (y > p) ? param : (1.0f - param);
    o.l = s * lc;
    o.g = -s * t;
    o.h = s * sech2;  // 1/cosh^2
  } else if constexpr (LOSS == SE_LOSS_BERNOULLI) {  // :293-318
    const float ye = 2.0f * y - 1.0f;
    const float z = 2.0f * ye * p;             // loss = log1pExp(-z)
    const float t = exp_neg_fast(-fabsf(z));   // in (0,1]
    const float inv = rcp_approx(1.0f + t);
    const float tinv = t * inv;
    const float sig_neg = (z >= 0.f) ? tinv : inv;  // sigma(-z) = 1/(1+e^z)
    const float sig_pos = (z >= 0.f) ? inv : tinv;  // sigma(z): formed directly, no 1-x cancellation
    o.l = fmaxf(-z, 0.f) + log1p_unit(t);
    o.g = -2.0f * ye * sig_neg;
    o.h = 4.0f * ye * ye * sig_neg * sig_pos;  // 4 e^z y^2 / (1+e^z)^2
  } else if constexpr (LOSS == SE_LOSS_EXPONENTIAL) {  // :265-291
    const float ye = 2.0f * y - 1.0f;
    const float e = exp_fast(-ye * p);
    o.l = e;
    o.g = -ye * e;
    o.h = ye * ye * e;
  } else {
    o.l = o.g = o.h = 0.f;
  }
  return o;
}

__host__ __device__ __forceinline__ bool loss_has_hessian(int loss) {
  return loss == SE_LOSS_SQUARED || loss == SE_LOSS_LOGCOSH || loss == SE_LOSS_SCALED_LOGCOSH ||
         loss == SE_LOSS_BERNOULLI || loss == SE_LOSS_EXPONENTIAL || loss == SE_LOSS_LOGLOSS;
}

}  // namespace se
