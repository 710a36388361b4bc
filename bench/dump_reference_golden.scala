// dump_reference_golden.scala — writes golden vectors FROM THE REAL REFERENCE (pierrenodet/spark-ensemble on Spark
// 3.3.1) for the hot path of this repository.  NOT runnable in this repository's image (no JVM); on any box with a
// JDK 8/11, Spark 3.3.1 and the spark-ensemble jar, ONE command makes the parity tests reference-pinned:
//
//   cd <checkout of pierrenodet/spark-ensemble>     # data/cpusmall/cpusmall.svm, data/adult/adult.svm live here
//   spark-shell --master 'local[*]' --jars core/target/scala-2.12/spark-ensemble_2.12-<version>.jar \
//       -i <this repo>/bench/dump_reference_golden.scala
//   cp reference_c1.json reference_c3.json <this repo>/tests/golden/
//
// tests/test_reference_golden.py consumes tests/golden/reference_*.json automatically when present (skipped
// otherwise) and replays every round through the C ABI on the GPU and through the oracle on the CPU:
//   * the base models are third party (Spark's DecisionTreeRegressor), so the dump carries their per-row training
//     predictions h_t (the "direction" of round t) — everything downstream of them IS the hot path of this repository;
//   * per round: the model weight learningRate * alpha_t (regression/GBMRegressor.scala:427,
//     classification/GBMClassifier.scala:432-436), the train loss of the updated F (boosting/GBMLoss.scala) and a
//     checksum of the running predictions (GBMRegressor.scala:434-442).
import java.io.PrintWriter
import org.apache.spark.ml.classification.{GBMClassificationModel, GBMClassifier}
import org.apache.spark.ml.linalg.{DenseVector, SparseVector, Vector, Vectors}
import org.apache.spark.ml.regression.{DecisionTreeRegressor, GBMRegressionModel, GBMRegressor}

def sliceVec(indices: Array[Int])(features: Vector): Vector = features match {  // ensemble/HasSubBag.scala:81-84
  case f: DenseVector => Vectors.dense(indices.map(f.apply))
  case f: SparseVector => f.slice(indices, true)
}
def arr(a: Array[Double]): String = a.map(v => java.lang.Double.toString(v)).mkString("[", ",", "]")

// ------------------------------------------------------------------ C1: GBMRegressor, cpusmall, squared, 20 rounds
{
  val df = spark.read.format("libsvm").load("data/cpusmall/cpusmall.svm").coalesce(1).cache()
  val rows = df.select("label", "features").collect().map(r => (r.getDouble(0), r.getAs[Vector](1)))
  val gbm = new GBMRegressor().setBaseLearner(new DecisionTreeRegressor().setMaxDepth(5)).setNumBaseLearners(20)
    .setLoss("squared")
  val model: GBMRegressionModel = gbm.fit(df)
  val y = rows.map(_._1)
  var F = rows.map { case (_, x) => model.init.predict(x) }
  val out = new PrintWriter("reference_c1.json")
  out.println("{\"generator\": \"bench/dump_reference_golden.scala\", \"spark\": \"" + spark.version + "\",")
  out.println(" \"config\": {\"estimator\": \"GBMRegressor\", \"data\": \"data/cpusmall/cpusmall.svm\", \"loss\": \"squared\", " +
    "\"numBaseLearners\": 20, \"learningRate\": " + gbm.getLearningRate + ", \"optimizedWeights\": " + gbm.getOptimizedWeights +
    ", \"tol\": " + gbm.getTol + ", \"maxIter\": " + gbm.getMaxIter + ", \"baseLearner\": \"DecisionTreeRegressor(maxDepth=5)\"},")
  out.println(" \"n\": " + y.length + ", \"init\": " + arr(F.take(1)) + ", \"rounds\": [")
  for (t <- 0 until model.numModels) {
    val h = rows.map { case (_, x) => model.models(t).predict(sliceVec(model.subspaces(t))(x)) }
    F = F.zip(h).map { case (f, hh) => f + model.weights(t) * hh }                       // GBMRegressor.scala:434-441
    val loss = y.zip(F).map { case (yy, f) => 0.5 * (yy - f) * (yy - f) }.sum / y.length  // GBMLoss.scala:129-137
    out.println("  {\"weight\": " + model.weights(t) + ", \"train_loss_after\": " + loss + ", \"prediction_checksum\": " +
      arr(Array(F.sum, F.map(v => v * v).sum)) + ", \"direction\": " + arr(h) + "}" + (if (t + 1 < model.numModels) "," else ""))
  }
  out.println(" ]}")
  out.close()
  println("wrote reference_c1.json")
}

// ------------------------------------------------------------------ C3 (scaled): GBMClassifier, adult (first 8000 rows), bernoulli
{
  val df = spark.read.format("libsvm").option("numFeatures", "123").load("data/adult/adult.svm").coalesce(1).limit(8000)
    .selectExpr("(label + 1) / 2 as label", "features").cache()                          // labels -1/+1 -> 0/1
  val rows = df.select("label", "features").collect().map(r => (r.getDouble(0), r.getAs[Vector](1)))
  val gbm = new GBMClassifier().setBaseLearner(new DecisionTreeRegressor().setMaxDepth(5)).setNumBaseLearners(10)
    .setLoss("bernoulli")
  val model: GBMClassificationModel = gbm.fit(df)
  val y = rows.map(_._1)
  // dim == 1 for bernoulli: raw = (-F, F) (classification/GBMClassifier.scala:583-584); F0 from the init model
  var F = rows.map { case (_, x) => model.init.predictRaw(x)(0) }
  val out = new PrintWriter("reference_c3.json")
  out.println("{\"generator\": \"bench/dump_reference_golden.scala\", \"spark\": \"" + spark.version + "\",")
  out.println(" \"config\": {\"estimator\": \"GBMClassifier\", \"data\": \"data/adult/adult.svm (first 8000 rows, labels (y+1)/2)\", " +
    "\"loss\": \"bernoulli\", \"numBaseLearners\": 10, \"learningRate\": " + gbm.getLearningRate + ", \"tol\": " + gbm.getTol +
    ", \"maxIter\": " + gbm.getMaxIter + ", \"initStrategy\": \"" + gbm.getInitStrategy + "\", \"baseLearner\": \"DecisionTreeRegressor(maxDepth=5)\"},")
  out.println(" \"n\": " + y.length + ", \"init\": " + arr(F.take(1)) + ", \"rounds\": [")
  for (t <- 0 until model.numModels) {
    val h = rows.map { case (_, x) => model.models(t)(0).predict(sliceVec(model.subspaces(t))(x)) }
    val w = model.weights(t)(0)
    F = F.zip(h).map { case (f, hh) => f + w * hh }                                      // GBMClassifier.scala:437-449
    val loss = y.zip(F).map { case (yy, f) =>                                             // BernoulliLoss, GBMLoss.scala:297-301
      val z = -2.0 * (2.0 * yy - 1.0) * f
      if (z > 0) z + math.log1p(math.exp(-z)) else math.log1p(math.exp(z)) }.sum / y.length
    out.println("  {\"weight\": " + w + ", \"train_loss_after\": " + loss + ", \"prediction_checksum\": " +
      arr(Array(F.sum, F.map(v => v * v).sum)) + ", \"direction\": " + arr(h) + "}" + (if (t + 1 < model.numModels) "," else ""))
  }
  out.println(" ]}")
  out.close()
  println("wrote reference_c3.json")
}
System.exit(0)
