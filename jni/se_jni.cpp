// se_jni.cpp — thin JNI shim over the C ABI of include/se_abi.h for the Scala side
// (scala/org/apache/spark/ml/se/SeNative.scala).  One JNI function per ABI entry point used by the
// reference's train()/predict() bodies; non-zero status becomes a RuntimeException
// (IllegalArgumentException for SE_ERR_ARG), mirroring the reference's error conventions
// (SURVEY.md §8b).  Host arrays are borrowed for the duration of the call
// (GetPrimitiveArrayCritical); device memory is owned by the se_ctx.
//
// Not compiled in this image: there is no JDK here (no <jni.h>).  Build where a JDK exists with
//   g++ -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude (continued)
//       jni/se_jni.cpp -Lspark_ensemble_b200/lib -lse_b200 -o libse_jni.so
#if defined(__has_include)
#if __has_include(<jni.h>)
#define SE_HAVE_JNI 1
#endif
#endif

#ifdef SE_HAVE_JNI
#include <jni.h>

#include <vector>

#include "../include/se_abi.h"

namespace {

inline se_ctx* H(jlong h) { return reinterpret_cast<se_ctx*>(h); }

bool raise(JNIEnv* env, se_ctx* ctx, int rc) {
  if (rc == SE_OK) return false;
  const char* cls = (rc == SE_ERR_ARG) ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException";
  env->ThrowNew(env->FindClass(cls), se_last_error(ctx));
  return true;
}

struct FloatPin {  // borrowed float[] for one call
  JNIEnv* env; jfloatArray arr; jfloat* p; jint mode;
  FloatPin(JNIEnv* e, jfloatArray a, jint m) : env(e), arr(a), mode(m) {
    p = a ? static_cast<jfloat*>(e->GetPrimitiveArrayCritical(a, nullptr)) : nullptr;
  }
  ~FloatPin() { if (p) env->ReleasePrimitiveArrayCritical(arr, p, mode); }
};
struct DoublePin {
  JNIEnv* env; jdoubleArray arr; jdouble* p; jint mode;
  DoublePin(JNIEnv* e, jdoubleArray a, jint m) : env(e), arr(a), mode(m) {
    p = a ? static_cast<jdouble*>(e->GetPrimitiveArrayCritical(a, nullptr)) : nullptr;
  }
  ~DoublePin() { if (p) env->ReleasePrimitiveArrayCritical(arr, p, mode); }
};

}  // namespace

#define SE_JNI(ret, name) extern "C" JNIEXPORT ret JNICALL Java_org_apache_spark_ml_se_SeNative_##name

SE_JNI(jlong, ctxCreate)(JNIEnv* env, jclass, jint device) {
  se_ctx* ctx = nullptr;
  if (raise(env, nullptr, se_ctx_create(device, &ctx))) return 0;
  return reinterpret_cast<jlong>(ctx);
}
SE_JNI(void, ctxDestroy)(JNIEnv*, jclass, jlong h) { se_ctx_destroy(H(h)); }

SE_JNI(jbyteArray, commUniqueId)(JNIEnv* env, jclass) {
  jbyte id[SE_COMM_ID_BYTES];
  if (raise(env, nullptr, se_comm_unique_id(id, SE_COMM_ID_BYTES))) return nullptr;
  jbyteArray out = env->NewByteArray(SE_COMM_ID_BYTES);
  env->SetByteArrayRegion(out, 0, SE_COMM_ID_BYTES, id);
  return out;
}
SE_JNI(void, commInit)(JNIEnv* env, jclass, jlong h, jint nranks, jint rank, jbyteArray id) {
  jbyte buf[SE_COMM_ID_BYTES] = {0};
  if (id) env->GetByteArrayRegion(id, 0, SE_COMM_ID_BYTES, buf);
  raise(env, H(h), se_comm_init(H(h), nranks, rank, id ? buf : nullptr, id ? SE_COMM_ID_BYTES : 0));
}

SE_JNI(void, upload)(JNIEnv* env, jclass, jlong h, jint slot, jfloatArray host, jlong count, jlong offset) {
  int rc;
  { FloatPin p(env, host, JNI_ABORT); rc = se_upload(H(h), slot, p.p, count, offset); }
  raise(env, H(h), rc);
}
SE_JNI(void, uploadF64)(JNIEnv* env, jclass, jlong h, jint slot, jdoubleArray host, jlong count, jlong offset) {
  int rc;
  { DoublePin p(env, host, JNI_ABORT); rc = se_upload_f64(H(h), slot, p.p, count, offset); }
  raise(env, H(h), rc);
}
SE_JNI(void, download)(JNIEnv* env, jclass, jlong h, jint slot, jfloatArray host, jlong count, jlong offset) {
  int rc;
  { FloatPin p(env, host, 0); rc = se_download(H(h), slot, p.p, count, offset); }
  raise(env, H(h), rc);
}
SE_JNI(void, fill)(JNIEnv* env, jclass, jlong h, jint slot, jfloat v, jlong count, jlong offset) {
  raise(env, H(h), se_fill(H(h), slot, v, count, offset));
}
SE_JNI(jdouble, slotSum)(JNIEnv* env, jclass, jlong h, jint slot, jlong count) {
  double s = 0.0;
  raise(env, H(h), se_slot_sum(H(h), slot, count, &s));
  return s;
}

SE_JNI(void, gbmConfigure)(JNIEnv* env, jclass, jlong h, jlong n, jlong nv, jint dim, jint loss, jdouble param,
                           jboolean hasWeights) {
  raise(env, H(h), se_gbm_configure(H(h), n, nv, dim, loss, param, hasWeights ? 1 : 0));
}
SE_JNI(void, gbmSetLossParam)(JNIEnv* env, jclass, jlong h, jdouble param) {
  raise(env, H(h), se_gbm_set_loss_param(H(h), param));
}
SE_JNI(void, gbmPseudoResiduals)(JNIEnv* env, jclass, jlong h, jboolean newton, jdoubleArray sumHess) {
  int rc;
  { DoublePin s(env, sumHess, 0); rc = se_gbm_pseudo_residuals(H(h), newton ? 1 : 0, s.p); }
  raise(env, H(h), rc);
}
// DiffFunction.calculate(alpha) => (loss, grad): what Breeze LBFGSB / commons-math3 Brent call per evaluation
SE_JNI(jdouble, gbmLinesearchEval)(JNIEnv* env, jclass, jlong h, jdoubleArray alpha, jdoubleArray grad) {
  double loss = 0.0;
  int rc;
  { DoublePin a(env, alpha, JNI_ABORT); DoublePin g(env, grad, 0); rc = se_gbm_linesearch_eval(H(h), a.p, &loss, g.p); }
  raise(env, H(h), rc);
  return loss;
}
SE_JNI(void, gbmLinesearchStats)(JNIEnv* env, jclass, jlong h, jdoubleArray stats4) {
  int rc;
  { DoublePin s(env, stats4, 0); rc = se_gbm_linesearch_stats(H(h), s.p); }
  raise(env, H(h), rc);
}
SE_JNI(jdouble, gbmUpdate)(JNIEnv* env, jclass, jlong h, jdoubleArray step, jint flags, jdoubleArray sumHess) {
  double loss = 0.0;
  int rc;
  { DoublePin s(env, step, JNI_ABORT); DoublePin sh(env, sumHess, 0); rc = se_gbm_update(H(h), s.p, flags, &loss, sh.p); }
  raise(env, H(h), rc);
  return loss;
}
SE_JNI(jdouble, gbmMeanLoss)(JNIEnv* env, jclass, jlong h, jint which) {
  double v = 0.0;
  raise(env, H(h), se_gbm_mean_loss(H(h), which, &v));
  return v;
}
SE_JNI(jdouble, gbmUpdateValidation)(JNIEnv* env, jclass, jlong h, jdoubleArray step) {
  double v = 0.0;
  int rc;
  { DoublePin s(env, step, JNI_ABORT); rc = se_gbm_update_validation(H(h), s.p, &v); }
  raise(env, H(h), rc);
  return v;
}

SE_JNI(void, boostConfigure)(JNIEnv* env, jclass, jlong h, jlong n, jint numClasses, jboolean real) {
  raise(env, H(h), se_boost_configure(H(h), n, numClasses, real ? 1 : 0));
}
SE_JNI(jdoubleArray, boostRealUpdate)(JNIEnv* env, jclass, jlong h, jdouble sumW) {
  double out[2] = {0, 0};
  if (raise(env, H(h), se_boost_real_update(H(h), sumW, &out[0], &out[1]))) return nullptr;
  jdoubleArray r = env->NewDoubleArray(2);
  env->SetDoubleArrayRegion(r, 0, 2, out);
  return r;  // (estimatorError, sumWeights')
}
SE_JNI(jdouble, boostDiscreteError)(JNIEnv* env, jclass, jlong h, jdouble sumW) {
  double e = 0.0;
  raise(env, H(h), se_boost_discrete_error(H(h), sumW, &e));
  return e;
}
SE_JNI(jdouble, boostDiscreteUpdate)(JNIEnv* env, jclass, jlong h, jdouble sumW, jdouble beta) {
  double s = 0.0;
  raise(env, H(h), se_boost_discrete_update(H(h), sumW, beta, &s));
  return s;
}

// ---- BoostingRegressor (AdaBoost.R2), quantiles, row sub-sampling, opt-in Newton line search ----------
SE_JNI(void, boostregConfigure)(JNIEnv* env, jclass, jlong h, jlong n) { raise(env, H(h), se_boostreg_configure(H(h), n)); }
SE_JNI(jdouble, boostregMaxError)(JNIEnv* env, jclass, jlong h) {
  double v = 0.0;
  raise(env, H(h), se_boostreg_max_error(H(h), &v));
  return v;
}
SE_JNI(jdouble, boostregError)(JNIEnv* env, jclass, jlong h, jdouble sumW, jint lossType, jdouble maxError) {
  double v = 0.0;
  raise(env, H(h), se_boostreg_error(H(h), sumW, lossType, maxError, &v));
  return v;
}
SE_JNI(jdouble, boostregUpdate)(JNIEnv* env, jclass, jlong h, jdouble sumW, jint lossType, jdouble maxError, jdouble beta) {
  double v = 0.0;
  raise(env, H(h), se_boostreg_update(H(h), sumW, lossType, maxError, beta, &v));
  return v;
}
SE_JNI(jdouble, quantile)(JNIEnv* env, jclass, jlong h, jint which, jint slot, jlong count, jdouble q) {
  double v = 0.0;
  raise(env, H(h), se_quantile(H(h), which, slot, count, q, &v));
  return v;
}
SE_JNI(void, gbmSetBag)(JNIEnv* env, jclass, jlong h, jboolean on) { raise(env, H(h), se_gbm_set_bag(H(h), on ? 1 : 0)); }
SE_JNI(jdoubleArray, gbmLinesearchNewton)(JNIEnv* env, jclass, jlong h, jdouble lo, jdouble hi, jdouble start, jdouble tol,
                                          jint maxEval) {
  double out[2] = {0, 0};
  int ne = 0;
  if (raise(env, H(h), se_gbm_linesearch_newton(H(h), lo, hi, start, tol, tol, maxEval, &out[0], &out[1], &ne))) return nullptr;
  jdoubleArray r = env->NewDoubleArray(2);
  env->SetDoubleArrayRegion(r, 0, 2, out);
  return r;  // (alpha, objective)
}

SE_JNI(void, aggConfigure)(JNIEnv* env, jclass, jlong h, jint kind, jint numModels, jint numClasses, jint dim,
                           jint loss, jlong n) {
  raise(env, H(h), se_agg_configure(H(h), kind, numModels, numClasses, dim, loss, n));
}
SE_JNI(void, aggRun)(JNIEnv* env, jclass, jlong h, jdoubleArray weights, jdoubleArray init) {
  int rc;
  { DoublePin w(env, weights, JNI_ABORT); DoublePin i(env, init, JNI_ABORT); rc = se_agg_run(H(h), w.p, i.p); }
  raise(env, H(h), rc);
}

#endif  // SE_HAVE_JNI
