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

// log(cosh(d)) without overflow and without cancellation near 0
__device__ __forceinline__ float logcoshf_(float d) {
  const float a = fabsf(d);
  if (a < 0.25f) {
    const float x2 = d * d;
    // x^2/2 - x^4/12 + x^6/45 - 17 x^8/2520
    return x2 * (0.5f + x2 * (-1.0f / 12.0f + x2 * (1.0f / 45.0f + x2 * (-17.0f / 2520.0f))));
  }
  return a + log1pf(expf(-2.0f * a)) - 0.69314718055994531f;
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
    const float t = tanhf(d);
    float s = 1.0f;
    if constexpr (LOSS == SE_LOSS_SCALED_LOGCOSH) s = (y > p) ? param : (1.0f - param);
    o.l = s * logcoshf_(d);
    o.g = -s * t;
    o.h = s * (1.0f - t * t);  // 1/cosh^2
  } else if constexpr (LOSS == SE_LOSS_BERNOULLI) {  // :293-318
    const float ye = 2.0f * y - 1.0f;
    const float z = 2.0f * ye * p;             // loss = log1pExp(-z)
    const float t = expf(-fabsf(z));           // in (0,1]
    const float inv = 1.0f / (1.0f + t);
    const float sig_neg = (z >= 0.f) ? t * inv : inv;  // sigma(-z) = 1/(1+e^z)
    const float sig_pos = (z >= 0.f) ? inv : t * inv;  // sigma(z): formed directly, no 1-x cancellation
    o.l = fmaxf(-z, 0.f) + log1pf(t);
    o.g = -2.0f * ye * sig_neg;
    o.h = 4.0f * ye * ye * sig_neg * sig_pos;  // 4 e^z y^2 / (1+e^z)^2
  } else if constexpr (LOSS == SE_LOSS_EXPONENTIAL) {  // :265-291
    const float ye = 2.0f * y - 1.0f;
    const float e = expf(-ye * p);
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
