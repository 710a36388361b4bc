// spark_baseline.scala — the reference's own CPU run of BASELINE config 1 (and scaled-down 2/3), to be executed
// wherever a JDK 8/11 + Spark 3.3.1 + the spark-ensemble jar exist.  NOT runnable in this repository's image
// (no JVM): shipped so the Spark local[*] number can be produced next to the GPU numbers of bench.py.
//
//   spark-shell --master 'local[*]' --jars spark-ensemble_2.12-<version>.jar -i bench/spark_baseline.scala
//
// Prints Runtime.availableProcessors, seconds per boosting round and rows/s (rows x rounds / wall time).
import org.apache.spark.ml.regression.{DecisionTreeRegressor, GBMRegressor}
import org.apache.spark.ml.classification.GBMClassifier
import org.apache.spark.ml.linalg.Vectors
import org.apache.spark.sql.functions._

val cores = Runtime.getRuntime.availableProcessors
println(s"host cores: $cores")

def timed[T](label: String, rows: Long, rounds: Int)(body: => T): T = {
  val t0 = System.nanoTime
  val out = body
  val s = (System.nanoTime - t0) / 1e9
  println(f"$label: $s%.2f s total, ${s / rounds}%.3f s/round, ${rows.toDouble * rounds / s}%.3e rows/s on $cores cores")
  out
}

// config 1: GBMRegressor on data/cpusmall, 20 rounds, squared loss
val cpusmall = spark.read.format("libsvm").load("data/cpusmall/cpusmall.svm").cache()
val n1 = cpusmall.count()
timed("C1 GBMRegressor cpusmall squared 20 rounds", n1, 20) {
  new GBMRegressor().setBaseLearner(new DecisionTreeRegressor().setMaxDepth(5)).setNumBaseLearners(20)
    .setLoss("squared").fit(cpusmall)
}

// scaled-down config 2: synthetic N x 64, squared loss (N chosen to fit the host)
val n2 = 2000000L
val d = 64
val synth = spark.range(n2).select(
  (rand(2) * 2 - 1).as("label"),
  udf((seed: Long) => { val r = new scala.util.Random(seed); Vectors.dense(Array.fill(d)(r.nextGaussian())) })
    .apply(col("id")).as("features")).cache()
synth.count()
timed(s"C2' GBMRegressor synthetic ${n2}x$d squared 10 rounds", n2, 10) {
  new GBMRegressor().setBaseLearner(new DecisionTreeRegressor().setMaxDepth(5)).setNumBaseLearners(10).fit(synth)
}

// scaled-down config 3: binary labels, bernoulli loss
val synthBin = synth.withColumn("label", when(col("label") > 0, 1.0).otherwise(0.0)).cache()
synthBin.count()
timed(s"C3' GBMClassifier synthetic ${n2}x$d bernoulli 10 rounds", n2, 10) {
  new GBMClassifier().setBaseLearner(new DecisionTreeRegressor().setMaxDepth(5)).setNumBaseLearners(10)
    .setLoss("bernoulli").fit(synthBin)
}
