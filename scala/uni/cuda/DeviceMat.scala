// DeviceMat.scala -- device-resident backing for `Mat[Double]` / `Mat[Float]`: the Scala twin of
// uni_b200/mat.py.  NOT compiled in the build image (no JVM).  It carries the reference's
// descriptor (Mat.scala:125-133) next to a device pointer, so `T`, `slice`, `broadcastTo` stay
// zero-copy, and forwards each operator of SURVEY.md section 8a to one libunimat call.
package uni.cuda

import java.lang.foreign.*
import java.lang.foreign.ValueLayout.*
import java.lang.ref.Cleaner
import UniMatLib.*

/** One device allocation, freed by a Cleaner when the last view dies (views retain their parent). */
final class DeviceBuffer(val ptr: MemorySegment, val bytes: Long):
  DeviceBuffer.cleaner.register(this, DeviceBuffer.Free(ptr))
object DeviceBuffer:
  private val cleaner = Cleaner.create()
  private final class Free(p: MemorySegment) extends Runnable:
    // runs on the Cleaner's thread: um_free queues the free on the allocating thread's stream and device (unimat.h)
    def run(): Unit =
      val rc = um_free.invoke(p).asInstanceOf[Int]
      if rc != 0 then System.err.println(s"libunimat: um_free failed ($rc): ${lastError()}")
  def alloc(bytes: Long): DeviceBuffer =
    val out = Arena.global().allocate(ADDRESS)
    check(um_malloc.invoke(math.max(bytes, 1L), out).asInstanceOf[Int])
    DeviceBuffer(out.get(ADDRESS, 0), bytes)

enum DType(val code: Int, val size: Int):
  case F64 extends DType(0, 8); case F32 extends DType(1, 4); case Bool extends DType(2, 1); case I32 extends DType(3, 4)

/** `MatData` on the device.  `view` is a 56-byte `um_view` segment owned by this object. */
final class DeviceMat private (val buf: DeviceBuffer, val view: MemorySegment, val dtype: DType):
  private inline def L(name: String) = view.get(JAVA_LONG, VIEW.byteOffset(MemoryLayout.PathElement.groupElement(name)))
  def rows: Int = L("rows").toInt
  def cols: Int = L("cols").toInt
  def rs: Long  = L("rs")
  def cs: Long  = L("cs")

  private def like(r: Long, c: Long, dt: DType = dtype): DeviceMat = DeviceMat.alloc(r, c, dt)
  private def derived(f: MemorySegment => Int): DeviceMat =
    val v = Arena.ofAuto().allocate(VIEW); check(f(v)); DeviceMat(buf, v, dtype)

  // ---- views: Mat.scala:198-199 (transposeView), :1905-1927 (slice), :1933-1953 (broadcastTo)
  def T: DeviceMat = derived(v => um_view_transpose.invoke(view, v).asInstanceOf[Int])
  def slice(r0: Int, r1: Int, c0: Int, c1: Int): DeviceMat =
    val need = Arena.ofAuto().allocate(JAVA_INT)
    val s = derived(v => um_view_slice.invoke(view, r0.toLong, r1.toLong, c0.toLong, c1.toLong, v, need).asInstanceOf[Int])
    if need.get(JAVA_INT, 0) != 0 then s.matCopy else s      // Internal.create's layout guard (Mat.scala:1865-1876)
  def broadcastTo(r: Int, c: Int): DeviceMat = derived(v => um_view_broadcast.invoke(view, r.toLong, c.toLong, v).asInstanceOf[Int])
  def matCopy: DeviceMat = { val o = like(rows, cols); check(um_copy.invoke(view, o.view).asInstanceOf[Int]); o }
  /** Row-major host copy (Mat.toArray): views are materialised first, then one D2H copy. */
  def toArray: Array[Double] =
    val c = if isContiguousView then this else matCopy
    val n = rows.toLong * cols.toLong
    val seg = Arena.ofAuto().allocate(n * 8)
    check(um_d2h.invoke(seg, MemorySegment.ofAddress(c.buf.ptr.address() + c.L("offset") * 8), n * 8).asInstanceOf[Int])
    seg.toArray(JAVA_DOUBLE)
  private def isContiguousView: Boolean =
    val o = Arena.ofAuto().allocate(JAVA_INT)
    check(um_view_is_contiguous.invoke(view, o).asInstanceOf[Int]); o.get(JAVA_INT, 0) != 0

  // ---- elementwise (Mat.scala:1989-2102, 2154-2209, 2514-2531, 4513-4588)
  private def bin(op: Int, b: DeviceMat): DeviceMat =
    val rc = Arena.ofAuto().allocate(JAVA_LONG, 2)
    check(um_broadcast_shape.invoke(view, b.view, rc, rc.asSlice(8)).asInstanceOf[Int])
    val o = like(rc.get(JAVA_LONG, 0), rc.get(JAVA_LONG, 8)); check(um_binary.invoke(op, view, b.view, o.view).asInstanceOf[Int]); o
  private def binS(op: Int, s: Double, left: Boolean = false): DeviceMat =
    val o = like(rows, cols); check(um_binary_scalar.invoke(op, view, s, if left then 1 else 0, o.view).asInstanceOf[Int]); o
  def +(b: DeviceMat) = bin(0, b); def -(b: DeviceMat) = bin(1, b); def *:*(b: DeviceMat) = bin(2, b); def /(b: DeviceMat) = bin(3, b)
  def +(s: Double) = binS(0, s); def -(s: Double) = binS(1, s); def *(s: Double) = binS(2, s); def /(s: Double) = binS(3, s)
  def unary_- : DeviceMat = un(0)
  def :+=(s: Double): DeviceMat = { check(um_binary_scalar.invoke(0, view, s, 0, view).asInstanceOf[Int]); this }
  def :-=(s: Double): DeviceMat = { check(um_binary_scalar.invoke(1, view, s, 0, view).asInstanceOf[Int]); this }
  def :*=(s: Double): DeviceMat = { check(um_binary_scalar.invoke(2, view, s, 0, view).asInstanceOf[Int]); this }
  def :/=(s: Double): DeviceMat = { check(um_binary_scalar.invoke(3, view, s, 0, view).asInstanceOf[Int]); this }
  def :+=(b: DeviceMat): DeviceMat =
    require(b.rows == rows && b.cols == cols, "in-place op requires equal shapes")            // Mat.scala:4565
    check(um_binary.invoke(0, view, b.view, view).asInstanceOf[Int]); this
  private def un(op: Int, out: DType = dtype): DeviceMat =
    val o = like(rows, cols, out); check(um_unary.invoke(op, view, o.view).asInstanceOf[Int]); o
  def sigmoid: DeviceMat = un(5, DType.F64)                                                   // MatMathOps.scala:97-122
  def relu: DeviceMat    = un(6)                                                              // MatMathOps.scala:125-146
  def softmax(axis: Int = 1): DeviceMat =                                                     // MatMathOps.scala:156-226
    val o = like(rows, cols, DType.F64); check(um_softmax.invoke(view, axis, o.view).asInstanceOf[Int]); o

  // ---- reductions (Mat.scala:2251-2279, 3414-3492, 4699-4739, 4784-4805)
  private def red(kind: Int): Double =
    val o = Arena.ofAuto().allocate(JAVA_DOUBLE); check(um_reduce_all.invoke(kind, view, o).asInstanceOf[Int]); o.get(JAVA_DOUBLE, 0)
  private def redAxis(kind: Int, axis: Int): DeviceMat =
    val o = if axis == 0 then like(1, cols) else like(rows, 1)
    check(um_reduce_axis.invoke(kind, view, axis, o.view).asInstanceOf[Int]); o
  def sum: Double = red(0); def mean: Double = red(1); def variance: Double = red(2); def std: Double = red(3)
  def min: Double = red(4); def max: Double = red(5)
  def sum(axis: Int) = redAxis(0, axis); def mean(axis: Int) = redAxis(1, axis); def std(axis: Int) = redAxis(3, axis)
  def argmax: (Int, Int) =
    val rc = Arena.ofAuto().allocate(JAVA_LONG, 2); check(um_arg_all.invoke(1, view, rc, rc.asSlice(8)).asInstanceOf[Int])
    (rc.get(JAVA_LONG, 0).toInt, rc.get(JAVA_LONG, 8).toInt)
  def idxmin(axis: Int): DeviceMat =                                                          // MatPandasOps.scala:20-41
    val o = if axis == 0 then like(1, cols, DType.I32) else like(rows, 1, DType.I32)
    check(um_idx_axis.invoke(0, view, axis, o.view).asInstanceOf[Int]); o
  def gt(s: Double): DeviceMat =                                                              // Mat.scala:3948-3972
    val o = like(rows, cols, DType.Bool); check(um_compare_scalar.invoke(0, view, s, o.view).asInstanceOf[Int]); o

  // ---- running folds, m(mask), where (Mat.scala:3635-3745, 4008-4021, 1305-1336)
  private def cum(kind: Int, axis: Int): DeviceMat =
    require(axis == 0 || axis == 1, s"axis must be 0 or 1, got $axis")
    val o = like(rows, cols); check(um_cumulative.invoke(kind, view, axis, o.view).asInstanceOf[Int]); o
  def cumsum(axis: Int): DeviceMat = cum(0, axis)
  def cummax(axis: Int): DeviceMat = cum(1, axis)
  def cummin(axis: Int): DeviceMat = cum(2, axis)
  def cumsum: DeviceMat =
    val o = like(1, rows * cols); check(um_cumulative.invoke(0, view, -1, o.view).asInstanceOf[Int]); o
  def apply(mask: DeviceMat): DeviceMat =
    require(mask.rows == rows && mask.cols == cols, s"Mask shape (${mask.rows}, ${mask.cols}) must match matrix shape ($rows, $cols)")
    val k = Arena.ofAuto().allocate(JAVA_LONG); check(um_mask_count.invoke(mask.view, k).asInstanceOf[Int])
    val o = like(1, k.get(JAVA_LONG, 0)); check(um_mask_select.invoke(view, mask.view, o.view).asInstanceOf[Int]); o

  // ---- linear algebra (Mat.scala:2815-2881, 2979-3006)
  def *@(b: DeviceMat): DeviceMat =
    require(cols == b.rows, s"matmul shape mismatch: ${rows}x$cols *@ ${b.rows}x${b.cols}")
    val o = like(rows, b.cols); check(um_matmul.invoke(view, b.view, o.view).asInstanceOf[Int]); o
  def inverse: DeviceMat = { val o = like(rows, cols); check(um_inverse.invoke(view, o.view).asInstanceOf[Int]); o }

object DeviceMat:
  def where(cond: DeviceMat, x: DeviceMat, y: DeviceMat): DeviceMat =                        // Mat.scala:1305-1321
    val o = alloc(x.rows, x.cols, x.dtype); check(um_where.invoke(cond.view, x.view, y.view, o.view).asInstanceOf[Int]); o
  def where(cond: DeviceMat, x: Double, y: Double): DeviceMat =                              // Mat.scala:1323-1336
    val o = alloc(cond.rows, cond.cols); check(um_where_scalar.invoke(cond.view, x, y, o.view).asInstanceOf[Int]); o
  def alloc(rows: Long, cols: Long, dt: DType = DType.F64): DeviceMat =
    val b = DeviceBuffer.alloc(rows * cols * dt.size)
    val v = Arena.ofAuto().allocate(VIEW)
    check(um_view_init.invoke(v, b.ptr, rows, cols, dt.code).asInstanceOf[Int])
    new DeviceMat(b, v, dt)
  /** Upload a host `MatD`'s row-major data (Mat.toArray) to the device. */
  def fromHost(data: Array[Double], rows: Int, cols: Int): DeviceMat =
    val m = alloc(rows, cols)
    val seg = MemorySegment.ofArray(data)
    val off = Arena.ofAuto().allocate(data.length.toLong * 8); off.copyFrom(seg)
    check(um_h2d.invoke(m.buf.ptr, off, data.length.toLong * 8).asInstanceOf[Int]); m

/** The global NumPy-compatible generator on the device (Mat.scala:1461-1530: setSeed / rand / randn / normal /
 *  uniform / randint over data/NumPyRNG.scala).  The 40-byte `um_rng` state sits where `globalRNG` sits today; every
 *  fill consumes the PCG64 stream exactly as the sequential loops do, so seeded programs print the same numbers. */
object DeviceRandom:
  @volatile private var g: MemorySegment = seeded(0L)                       // lazy val defaultRNG = new NumPyRNG(0)
  private def seeded(seed: Long): MemorySegment =
    val s = Arena.ofAuto().allocate(RNG); check(um_rng_seed.invoke(s, seed).asInstanceOf[Int]); s   // seed < 0 -> IllegalArgumentException
  def setSeed(seed: Long): Unit = g = seeded(seed)
  def rand(rows: Int, cols: Int): DeviceMat =
    val o = DeviceMat.alloc(rows, cols); check(um_rng_rand.invoke(g, o.view).asInstanceOf[Int]); o
  def randn(rows: Int, cols: Int = 1): DeviceMat =
    val o = DeviceMat.alloc(rows, cols); check(um_rng_randn.invoke(g, o.view).asInstanceOf[Int]); o
  def normal(mean: Double, std: Double, rows: Int, cols: Int): DeviceMat =
    val o = DeviceMat.alloc(rows, cols); check(um_rng_normal.invoke(g, mean, std, o.view).asInstanceOf[Int]); o
  def uniform(low: Double, high: Double, rows: Int, cols: Int): DeviceMat =
    val o = DeviceMat.alloc(rows, cols); check(um_rng_uniform.invoke(g, low, high, o.view).asInstanceOf[Int]); o
  def randint(low: Int, high: Int, rows: Int, cols: Int): DeviceMat =
    val o = DeviceMat.alloc(rows, cols, DType.I32); check(um_rng_randint.invoke(g, low, high, o.view).asInstanceOf[Int]); o

/** `path.loadMatD` / `m.saveCSV` (ext/PathExts.scala:583-590, io/FileOps.scala:139-188, Mat.scala:90-116). */
object DeviceCsv:
  def loadSmartD(path: String, delimiter: Char = 0): (Vector[String], DeviceMat) =
    val a = Arena.ofAuto()
    val h = a.allocate(JAVA_LONG); val r = a.allocate(JAVA_LONG); val c = a.allocate(JAVA_LONG); val hd = a.allocate(JAVA_INT)
    check(um_csv_open.invoke(a.allocateFrom(path), delimiter.toInt, h, r, c, hd).asInstanceOf[Int])
    val handle = h.get(JAVA_LONG, 0)
    try
      val m = DeviceMat.alloc(r.get(JAVA_LONG, 0), c.get(JAVA_LONG, 0))
      check(um_csv_upload.invoke(handle, m.view).asInstanceOf[Int])
      val buf = a.allocate(4096)
      val headers = if hd.get(JAVA_INT, 0) == 0 then Vector.empty else Vector.tabulate(m.cols.toInt) { j =>
        check(um_csv_header.invoke(handle, j.toLong, buf, 4096L).asInstanceOf[Int]); buf.getString(0) }
      (headers, m)
    finally um_csv_close.invoke(handle)
  def loadMatD(path: String): DeviceMat = loadSmartD(path)._2
  def saveCSV(m: DeviceMat, path: String, sep: String = ",", nanAs: String = "NaN"): Unit =
    val a = Arena.ofAuto()
    check(um_csv_save.invoke(m.view, a.allocateFrom(path), a.allocateFrom(sep), a.allocateFrom(nanAs)).asInstanceOf[Int])

/** `uni.stats.Tprf3` on the device (Tprf3.scala:199-289, 883-943, 1056-1250): the Scala twin of uni_b200/tprf3.py. */
object DeviceTprf3:
  /** Tprf3Result (Tprf3.scala:243-252): every field of the reference's record; `alpha` only where the reference has it. */
  final case class Result(forecasts: DeviceMat, residuals: DeviceMat, rSquared: Double, alpha: Option[DeviceMat],
                          phi: DeviceMat, phiHat: DeviceMat, sigma: DeviceMat, beta3: DeviceMat, xstd: DeviceMat)
  /** estimate3prf's out-of-sample record: T-long host vectors (NaN where a window has no forecast), R^2, ENC-NEW. */
  final case class OosResult(forecasts: Array[Double], rollfore: Array[Double], rSquared: Double, encNew: Double)

  /** nTotal > 0 declares X this rank's column shard of an nTotal-column panel (um_comm_init done): one all-reduce per fit. */
  private def fit(mode: Int, y: DeviceMat, X: DeviceMat, Z: DeviceMat, nTotal: Long = 0L): Result =
    val (t, n, l) = (X.rows.toLong, X.cols.toLong, Z.cols.toLong)
    val f = DeviceMat.alloc(t, 1); val r = DeviceMat.alloc(t, 1); val a = DeviceMat.alloc(n, 1)
    val phi = DeviceMat.alloc(n, l); val phiHat = DeviceMat.alloc(l + 1, n); val sigma = DeviceMat.alloc(t, l)
    val beta3 = DeviceMat.alloc(l + 1, 1); val xstd = DeviceMat.alloc(1, n)
    val out = Arena.ofAuto().allocate(TPRF_OUT)
    out.set(ADDRESS, 0, f.buf.ptr); out.set(ADDRESS, 8, r.buf.ptr)
    if mode != 1 then out.set(ADDRESS, 16, a.buf.ptr)            // t3prf has no alpha (Tprf3.scala:260-289)
    out.set(ADDRESS, 24, phi.buf.ptr); out.set(ADDRESS, 32, phiHat.buf.ptr)
    out.set(ADDRESS, 40, sigma.buf.ptr); out.set(ADDRESS, 48, beta3.buf.ptr); out.set(ADDRESS, 56, xstd.buf.ptr)
    check(um_tprf_fit.invoke(mode, y.view, X.view, Z.view, nTotal, out).asInstanceOf[Int])   // status 2 -> ArithmeticException
    Result(f, r, out.get(JAVA_DOUBLE, 64), if mode != 1 then Some(a) else None, phi, phiHat, sigma, beta3, xstd)
  def tprfClosedForm(y: DeviceMat, X: DeviceMat, Z: DeviceMat, nTotal: Long = 0L): Result = fit(0, y, X, Z, nTotal)
  def t3prf(y: DeviceMat, X: DeviceMat, Z: DeviceMat, nTotal: Long = 0L): Result          = fit(1, y, X, Z, nTotal)
  def estimate3prfIsFull(y: DeviceMat, X: DeviceMat, Z: DeviceMat, nTotal: Long = 0L): Result = fit(2, y, X, Z, nTotal)

  /** estimate3prf(..., "OOS Recursive" | "OOS Cross Val" | "OOS Rolling") (Tprf3.scala:1112-1199): the window table the
   *  reference's `.par` loop walks, then ONE um_tprf_oos_windows call (every window's standardisation, three passes --
   *  automatic proxies rebuilt per window when `proxies` is Left(L) -- and held-out forecast in one launch), then the
   *  point estimates over the rows that have a forecast (:1207, 1226-1238, encnew :854-860). */
  def estimate3prfOos(y: DeviceMat, X: DeviceMat, proxies: Either[Int, DeviceMat], procedure: String,
                      window: (Int, Int) = (0, 1), mintrain: (Int, Int) = (-1, 0), rollwin: (Int, Int, Int) = (30, 20, 0)): OosResult =
    val t = y.rows
    val auto = proxies.isLeft
    val nAuto = proxies.left.getOrElse(0)
    val mt = if mintrain._1 < 0 then (t / 2, mintrain._2) else mintrain
    val (win, minNoNa, gap) = rollwin
    // (held-out row, lo, hi, kind, minObs): kind 0 keeps rows [lo, hi), kind 1 drops them
    val (ts, lo, hi, kind, minObs) = procedure match
      case "OOS Cross Val" =>
        val ts = (0 until t).toArray
        (ts, ts.map(i => math.max(i - window._1, 0)), ts.map(i => math.min(i - window._1 + window._2, t)), 1, 10)
      case "OOS Recursive" =>
        val ts = (mt._1 + 1 + mt._2 until t).toArray
        (ts, ts.map(_ => 0), ts.map(i => i - 1 - mt._2), 0, if auto then mt._1 else 10)
      case "OOS Rolling" =>
        val ts = (win + 1 + gap until t).toArray
        (ts, ts.map(i => math.max(i - win - gap, 0)), ts.map(i => math.min(i - 1 - gap, t)), 0, if auto then minNoNa else 10)
      case other => throw IllegalArgumentException(s"Unknown procedure '$other'. Choose: 'IS Full', 'OOS Recursive', 'OOS Cross Val', 'OOS Rolling'")
    val fore = Array.fill(t)(Double.NaN); val roll = Array.fill(t)(Double.NaN)
    if ts.nonEmpty then
      val a = Arena.ofAuto()
      val seg = (v: Array[Int]) => a.allocateFrom(JAVA_INT, v*)
      val fOut = a.allocate(JAVA_DOUBLE, ts.length.toLong); val rOut = a.allocate(JAVA_DOUBLE, ts.length.toLong)
      val hi2 = hi.zip(lo).map((h, l) => math.max(h, l))
      val xc = if X.cs == 1 || X.cols == 1 then X else X.matCopy            // the RAW panel: the call standardises it itself
      val z = proxies.fold(_ => MemorySegment.NULL, _.view)
      check(um_tprf_oos_windows.invoke(y.view, xc.view, z, nAuto, kind, seg(lo), seg(hi2), seg(ts), ts.length.toLong,
                                       minObs, 1, fOut, rOut).asInstanceOf[Int])
      for (i, k) <- ts.zipWithIndex do
        fore(i) = fOut.getAtIndex(JAVA_DOUBLE, k.toLong); roll(i) = rOut.getAtIndex(JAVA_DOUBLE, k.toLong)
    // R^2 with the rolling-mean denominator and ENC-NEW, summed in the reference's order
    val yh = y.toArray
    var see = 0.0; var sst = 0.0; var enc = 0.0; var cnt = 0
    var i = 0
    while i < t do
      val e = yh(i) - fore(i)
      if !e.isNaN then
        see += e * e; val d = yh(i) - roll(i); sst += d * d; enc += roll(i) * roll(i) - roll(i) * e; cnt += 1
      i += 1
    val r2 = if cnt > 0 && sst != 0.0 then 1.0 - see / sst else Double.NaN
    OosResult(fore, roll, r2, if procedure == "OOS Recursive" && cnt > 0 then cnt.toDouble * enc / see else Double.NaN)
