// se_brent.cu — the squared-loss line search on the device (regression/GBMRegressor.scala:408-421): Brent over the
// exact parabola (s0 - 2αs1 + α²s2)/(2Σw) of the three sufficient statistics, one thread.  The statistics kernel,
// this kernel and the fused update then run back to back on the stream — no host round trip inside the round.
//
// This translation unit is compiled with -fmad=false: every multiply and add rounds separately, exactly like the
// host build of the same template (se_brent.h), so α, the objective and the evaluation count equal the host line
// search bit for bit (tests/test_gpu_parity.py::test_device_brent_matches_host_brent).
#include "se_brent.h"
#include "se_kernels.h"

namespace se {

namespace {

__global__ void brent_parabola_kernel(const double* __restrict__ stats, double wsum, double lo, double hi, double start,
                                      double rel, double abs_tol, int max_eval, double* __restrict__ out_dev,
                                      double* __restrict__ out_host) {
  const BrentParabola f{stats[0], stats[1], stats[2], wsum};
  double x = 0.0, fx = 0.0;
  int evals = 0;
  const int rc = brent_core(f, lo, hi, start, rel, abs_tol, max_eval, &x, &fx, &evals);
  const double ne = (rc == kBrentOk) ? (double)evals : -(double)evals;  // negative: MaxEval exceeded
  out_dev[0] = x, out_dev[1] = fx, out_dev[2] = ne;
  if (out_host != nullptr) {
    out_host[0] = x, out_host[1] = fx, out_host[2] = ne;
    __threadfence_system();
  }
}

}  // namespace

cudaError_t launch_brent_parabola(const double* stats, double wsum, double lo, double hi, double start, double rel,
                                  double abs_tol, int max_eval, double* out_dev, double* out_host, cudaStream_t st) {
  brent_parabola_kernel<<<1, 1, 0, st>>>(stats, wsum, lo, hi, start, rel, abs_tol, max_eval, out_dev, out_host);
  return cudaGetLastError();
}

}  // namespace se
