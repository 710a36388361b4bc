/*
 * NativeLineSearch.scala — the objective the reference's optimisers see, backed by the GPU.
 *
 * The reference builds `new RDDLossFunction(instances, new GBMLossAggregator(loss)(_), None, depth)` and hands
 * it (through CachedDiffFunction) to commons-math3 Brent (regression/GBMRegressor.scala:408-421) or Breeze
 * LBFGSB (classification/GBMClassifier.scala:424-427).  This class is the replacement DiffFunction: one call
 * = one streaming pass of the K2 kernel over the HBM-resident (y, F, h) shard(s) + one tiny allreduce, and it
 * returns exactly Spark's (lossSum/weightSum, gradientSum/weightSum), including the `dim`-times loss quirk
 * (boosting/GBMLoss.scala:60-64).  The optimisers themselves are untouched.
 */
package org.apache.spark.ml.se

import breeze.linalg.{DenseVector => BDV}
import breeze.optimize.DiffFunction

class NativeLineSearch(ctx: Long, dim: Int) extends DiffFunction[BDV[Double]] {
  override def calculate(alpha: BDV[Double]): (Double, BDV[Double]) = {
    val grad = new Array[Double](dim)
    val loss = SeNative.gbmLinesearchEval(ctx, alpha.toArray, grad)
    (loss, new BDV(grad))
  }
}
