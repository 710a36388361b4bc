/*
 * SeNative.scala — JVM side of the drop-in boundary: @native bindings of jni/se_jni.cpp, which forwards
 * 1:1 to the C ABI of include/se_abi.h (libse_b200.so, sm_100a kernels).
 *
 * Not compiled in this repository's image (no JDK/scalac/sbt here); see INTEGRATION.md for how the
 * reference's train()/predict() bodies call these in place of their per-row RDD closures.
 */
package org.apache.spark.ml.se

object SeNative {
  System.loadLibrary("se_jni") // links libse_b200.so

  // enum se_slot / se_loss / se_agg_kind / update flags (include/se_abi.h)
  object Slot { val Y = 0; val W = 1; val F = 2; val H = 3; val R = 4; val WOUT = 5; val VY = 6; val VF = 7
    val VH = 8; val BW = 9; val PROBA = 10; val PRED = 11; val P = 12; val RAW = 13; val PROB = 14
    val LABEL = 15; val X = 16; val VX = 17; val BAG = 18 }
  object Loss { val Squared = 0; val Absolute = 1; val Huber = 2; val Quantile = 3; val LogCosh = 4
    val ScaledLogCosh = 5; val Bernoulli = 6; val Exponential = 7; val LogLoss = 8 }
  object Upd { val Residual = 1; val Newton = 2; val Loss = 4 }

  @native def ctxCreate(device: Int): Long
  @native def ctxDestroy(ctx: Long): Unit
  @native def commUniqueId(): Array[Byte]
  @native def commInit(ctx: Long, nranks: Int, rank: Int, id: Array[Byte]): Unit

  @native def upload(ctx: Long, slot: Int, host: Array[Float], count: Long, offset: Long): Unit
  @native def uploadF64(ctx: Long, slot: Int, host: Array[Double], count: Long, offset: Long): Unit
  @native def download(ctx: Long, slot: Int, host: Array[Float], count: Long, offset: Long): Unit
  @native def fill(ctx: Long, slot: Int, value: Float, count: Long, offset: Long): Unit
  @native def slotSum(ctx: Long, slot: Int, count: Long): Double

  @native def gbmConfigure(ctx: Long, n: Long, nValid: Long, dim: Int, loss: Int, param: Double,
      hasWeights: Boolean): Unit
  @native def gbmSetLossParam(ctx: Long, param: Double): Unit
  @native def gbmPseudoResiduals(ctx: Long, newton: Boolean, sumHess: Array[Double]): Unit
  @native def gbmLinesearchEval(ctx: Long, alpha: Array[Double], grad: Array[Double]): Double
  @native def gbmLinesearchStats(ctx: Long, stats4: Array[Double]): Unit
  @native def gbmUpdate(ctx: Long, step: Array[Double], flags: Int, sumHess: Array[Double]): Double
  @native def gbmMeanLoss(ctx: Long, which: Int): Double
  @native def gbmUpdateValidation(ctx: Long, step: Array[Double]): Double

  @native def boostConfigure(ctx: Long, n: Long, numClasses: Int, real: Boolean): Unit
  @native def boostRealUpdate(ctx: Long, sumWeights: Double): Array[Double] // (estimatorError, sumWeights')
  @native def boostDiscreteError(ctx: Long, sumWeights: Double): Double
  @native def boostDiscreteUpdate(ctx: Long, sumWeights: Double, beta: Double): Double

  // BoostingRegressor (AdaBoost.R2): lossType 0 exponential, 1 linear, 2 squared
  @native def boostregConfigure(ctx: Long, n: Long): Unit
  @native def boostregMaxError(ctx: Long): Double
  @native def boostregError(ctx: Long, sumWeights: Double, lossType: Int, maxError: Double): Double
  @native def boostregUpdate(ctx: Long, sumWeights: Double, lossType: Int, maxError: Double, beta: Double): Double
  // exact quantile (which = 0: slot values, 1: |y - F|), row sub-sampling multiplicities, opt-in Newton line search
  @native def quantile(ctx: Long, which: Int, slot: Int, count: Long, q: Double): Double
  @native def gbmSetBag(ctx: Long, on: Boolean): Unit
  @native def gbmLinesearchNewton(ctx: Long, lo: Double, hi: Double, start: Double, tol: Double, maxEval: Int): Array[Double]

  @native def aggConfigure(ctx: Long, kind: Int, numModels: Int, numClasses: Int, dim: Int, loss: Int,
      n: Long): Unit
  @native def aggRun(ctx: Long, weights: Array[Double], init: Array[Double]): Unit
}
