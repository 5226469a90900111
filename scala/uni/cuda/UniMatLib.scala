// UniMatLib.scala -- Panama (java.lang.foreign, JDK >= 22) downcall bindings for libunimat.so.
//
// NOT compiled in the build image (no JVM there, SURVEY.md section 8c); this is the binding a uni
// maintainer adds as a separate sbt module (the core library targets -source 17, build.sbt:6).
// Every handle below corresponds 1:1 to a prototype in include/unimat.h; the Python ctypes twin
// (uni_b200/_lib.py) binds the same symbols with the same struct layout and IS exercised by the
// test-suite, so the ABI described here is the tested one.
package uni.cuda

import java.lang.foreign.*
import java.lang.foreign.ValueLayout.*
import java.lang.invoke.MethodHandle

object UniMatLib:
  private val linker = Linker.nativeLinker()
  private val arena  = Arena.global()
  private val lookup: SymbolLookup =
    SymbolLookup.libraryLookup(sys.props.getOrElse("uni.mat.cuda.lib", "libunimat.so"), arena)

  // struct um_view { void* data; int64 offset, rows, cols, rs, cs; int32 dtype, transposed; }  (56 bytes)
  val VIEW: StructLayout = MemoryLayout.structLayout(
    ADDRESS.withName("data"), JAVA_LONG.withName("offset"), JAVA_LONG.withName("rows"), JAVA_LONG.withName("cols"),
    JAVA_LONG.withName("rs"), JAVA_LONG.withName("cs"), JAVA_INT.withName("dtype"), JAVA_INT.withName("transposed"))

  // struct um_tprf_out { 8 x double* ; double r_squared; int32 singular; (+4 pad) }  (80 bytes)
  val TPRF_OUT: StructLayout = MemoryLayout.structLayout(
    ADDRESS.withName("forecasts"), ADDRESS.withName("residuals"), ADDRESS.withName("alpha"), ADDRESS.withName("phi"),
    ADDRESS.withName("phi_hat"), ADDRESS.withName("sigma"), ADDRESS.withName("beta3"), ADDRESS.withName("xstd"),
    JAVA_DOUBLE.withName("r_squared"), JAVA_INT.withName("singular"), MemoryLayout.paddingLayout(4))

  private def h(name: String, fd: FunctionDescriptor): MethodHandle =
    linker.downcallHandle(lookup.find(name).orElseThrow(() => UnsatisfiedLinkError(name)), fd)
  private def I(args: MemoryLayout*) = FunctionDescriptor.of(JAVA_INT, args*)

  // ---- lifecycle / memory
  val um_init        = h("um_init", I(JAVA_INT))
  val um_sync        = h("um_sync", I())
  val um_last_error  = h("um_last_error", FunctionDescriptor.of(ADDRESS))
  val um_malloc      = h("um_malloc", I(JAVA_LONG, ADDRESS))
  val um_free        = h("um_free", I(ADDRESS))
  val um_malloc_host = h("um_malloc_host", I(JAVA_LONG, ADDRESS))
  val um_free_host   = h("um_free_host", I(ADDRESS))
  val um_h2d         = h("um_h2d", I(ADDRESS, ADDRESS, JAVA_LONG))
  val um_d2h         = h("um_d2h", I(ADDRESS, ADDRESS, JAVA_LONG))
  // ---- views (host-side descriptor arithmetic incl. the isWeirdLayout guard)
  val um_view_init      = h("um_view_init", I(ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT))
  val um_view_transpose = h("um_view_transpose", I(ADDRESS, ADDRESS))
  val um_view_slice     = h("um_view_slice", I(ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS))
  val um_view_broadcast = h("um_view_broadcast", I(ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS))
  val um_broadcast_shape = h("um_broadcast_shape", I(ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  val um_copy = h("um_copy", I(ADDRESS, ADDRESS))
  val um_fill = h("um_fill", I(ADDRESS, JAVA_DOUBLE))
  val um_get  = h("um_get", I(ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS))
  val um_set  = h("um_set", I(ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_DOUBLE))
  // ---- elementwise / activations / masks
  val um_binary         = h("um_binary", I(JAVA_INT, ADDRESS, ADDRESS, ADDRESS))
  val um_binary_scalar  = h("um_binary_scalar", I(JAVA_INT, ADDRESS, JAVA_DOUBLE, JAVA_INT, ADDRESS))
  val um_unary          = h("um_unary", I(JAVA_INT, ADDRESS, ADDRESS))
  val um_softmax        = h("um_softmax", I(ADDRESS, JAVA_INT, ADDRESS))
  val um_softmax_set_axis0_path = h("um_softmax_set_axis0_path", I(JAVA_INT))
  val um_compare_scalar = h("um_compare_scalar", I(JAVA_INT, ADDRESS, JAVA_DOUBLE, ADDRESS))
  val um_compare        = h("um_compare", I(JAVA_INT, ADDRESS, ADDRESS, ADDRESS))
  val um_classify       = h("um_classify", I(JAVA_INT, ADDRESS, ADDRESS))
  val um_mask_logic     = h("um_mask_logic", I(JAVA_INT, ADDRESS, ADDRESS, ADDRESS))
  val um_mask_reduce    = h("um_mask_reduce", I(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS))
  // ---- reductions
  val um_reduce_all  = h("um_reduce_all", I(JAVA_INT, ADDRESS, ADDRESS))
  val um_reduce_axis = h("um_reduce_axis", I(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS))
  val um_arg_all     = h("um_arg_all", I(JAVA_INT, ADDRESS, ADDRESS, ADDRESS))
  val um_idx_axis    = h("um_idx_axis", I(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS))
  // ---- linear algebra
  val um_cumulative  = h("um_cumulative", I(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS))
  val um_where       = h("um_where", I(ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  val um_where_scalar = h("um_where_scalar", I(ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS))
  val um_mask_count  = h("um_mask_count", I(ADDRESS, ADDRESS))
  val um_mask_select = h("um_mask_select", I(ADDRESS, ADDRESS, ADDRESS))
  val um_matmul  = h("um_matmul", I(ADDRESS, ADDRESS, ADDRESS))
  val um_matmul_set_f32_path = h("um_matmul_set_f32_path", I(JAVA_INT))
  val um_matmul_set_f64_tile = h("um_matmul_set_f64_tile", I(JAVA_INT))
  val um_inverse = h("um_inverse", I(ADDRESS, ADDRESS))
  val um_solve   = h("um_solve", I(ADDRESS, ADDRESS, ADDRESS))
  // ---- 3PRF + multi-GPU
  val um_tprf_fit = h("um_tprf_fit", I(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS))
  val um_tprf_fit_host = h("um_tprf_fit_host",
    I(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS))
  val um_tprf_fit_batched = h("um_tprf_fit_batched",
    I(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS))
  // which sweep the fits use (0 auto, 1 two-sweep, 2 cluster sweep, 3 grid sweep) and its shape on this device
  val um_tprf_set_path   = h("um_tprf_set_path", I(JAVA_INT))
  val um_tprf_oos_forecast = h("um_tprf_oos_forecast", I(ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  val um_tprf_fused_info = h("um_tprf_fused_info", I(JAVA_LONG, ADDRESS, ADDRESS, ADDRESS))
  val um_comm_unique_id = h("um_comm_unique_id", I(ADDRESS))
  val um_comm_init      = h("um_comm_init", I(JAVA_INT, JAVA_INT, ADDRESS))
  val um_comm_destroy   = h("um_comm_destroy", I())

  // ---- NumPy-compatible PCG64 generator (data/NumPyRNG.scala, Mat.scala:1461-1530) and CSV ingest / egress
  // struct um_rng { uint64 state_hi, state_lo, inc_hi, inc_lo; int32 has_uint32; uint32 uinteger; }  (40 bytes)
  val RNG: StructLayout = MemoryLayout.structLayout(
    JAVA_LONG.withName("state_hi"), JAVA_LONG.withName("state_lo"), JAVA_LONG.withName("inc_hi"), JAVA_LONG.withName("inc_lo"),
    JAVA_INT.withName("has_uint32"), JAVA_INT.withName("uinteger"))
  val um_rng_seed     = h("um_rng_seed", I(ADDRESS, JAVA_LONG))
  val um_rng_advance  = h("um_rng_advance", I(ADDRESS, JAVA_LONG))
  val um_rng_next_u64 = h("um_rng_next_u64", I(ADDRESS, ADDRESS))
  val um_rng_rand     = h("um_rng_rand", I(ADDRESS, ADDRESS))
  val um_rng_uniform  = h("um_rng_uniform", I(ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS))
  val um_rng_randn    = h("um_rng_randn", I(ADDRESS, ADDRESS))
  val um_rng_normal   = h("um_rng_normal", I(ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS))
  val um_rng_randint  = h("um_rng_randint", I(ADDRESS, JAVA_INT, JAVA_INT, ADDRESS))
  val um_csv_open     = h("um_csv_open", I(ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  val um_csv_header   = h("um_csv_header", I(JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_LONG))
  val um_csv_upload   = h("um_csv_upload", I(JAVA_LONG, ADDRESS))
  val um_csv_close    = h("um_csv_close", I(JAVA_LONG))
  val um_csv_save     = h("um_csv_save", I(ADDRESS, ADDRESS, ADDRESS, ADDRESS))

  // ---- the rest of include/unimat.h (tests/test_abi_cpu.py::test_scala_binding_matches_header keeps this file and
  //      the header in step: every prototype has exactly one handle with the same argument layouts)
  // device / lifecycle / raw memory
  val um_version      = h("um_version", I())
  val um_device_count = h("um_device_count", I(ADDRESS))
  val um_shutdown     = h("um_shutdown", I())
  val um_device_props = h("um_device_props", I(ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  val um_d2d          = h("um_d2d", I(ADDRESS, ADDRESS, JAVA_LONG))
  val um_memset       = h("um_memset", I(ADDRESS, JAVA_INT, JAVA_LONG))
  val um_h2d_2d       = h("um_h2d_2d", I(ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG))
  // measurement hooks (apps/Tprf3Bench.scala's stopwatch becomes device events)
  val um_launch_count = h("um_launch_count", I(ADDRESS))
  val um_timer_start  = h("um_timer_start", I())
  val um_timer_stop   = h("um_timer_stop", I(ADDRESS))
  val um_prof_enable  = h("um_prof_enable", I(JAVA_INT))
  val um_prof_reset   = h("um_prof_reset", I())
  val um_prof_read    = h("um_prof_read", I(JAVA_INT, ADDRESS, ADDRESS))
  val um_tprf_work    = h("um_tprf_work", I(JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS))
  val um_tprf_debug_stamps = h("um_tprf_debug_stamps", I(ADDRESS))
  // layout queries (Mat.scala:1865-1876 isWeirdLayout, isContiguous), factories (MatD.eye, counter-based fills)
  val um_view_is_weird      = h("um_view_is_weird", I(ADDRESS, ADDRESS))
  val um_view_is_contiguous = h("um_view_is_contiguous", I(ADDRESS, ADDRESS))
  val um_eye   = h("um_eye", I(ADDRESS))
  val um_randn = h("um_randn", I(ADDRESS, JAVA_LONG))
  val um_rand  = h("um_rand", I(ADDRESS, JAVA_LONG))
  // batched small inverse (the nanOls passes of estimate3prf(pls = true)) and the out-of-sample procedures in one launch
  val um_inverse_batched  = h("um_inverse_batched", I(ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_LONG))
  val um_tprf_oos_windows = h("um_tprf_oos_windows",
    I(ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS))
  // communicator queries and the raw exchange
  val um_comm_info              = h("um_comm_info", I(ADDRESS, ADDRESS))
  val um_comm_transport         = h("um_comm_transport", I(ADDRESS))
  val um_comm_allreduce_sum_f64 = h("um_comm_allreduce_sum_f64", I(ADDRESS, JAVA_LONG))
  val um_comm_barrier           = h("um_comm_barrier", I())
  // CSV cells on the host (FileOps.scala:139-188, Mat.scala:90-116)
  val um_csv_read_host  = h("um_csv_read_host", I(JAVA_LONG, ADDRESS, JAVA_LONG))
  val um_csv_format_f64 = h("um_csv_format_f64", I(JAVA_DOUBLE, ADDRESS, ADDRESS, JAVA_LONG))
  val um_csv_format_f32 = h("um_csv_format_f32", I(JAVA_FLOAT, ADDRESS, ADDRESS, JAVA_LONG))
  val um_csv_parse_cell = h("um_csv_parse_cell", I(ADDRESS, ADDRESS))

  // status -> the exception the reference throws for the same condition
  def check(status: Int): Unit =
    if status != 0 then
      val msg = um_last_error.invoke().asInstanceOf[MemorySegment].reinterpret(512).getString(0)
      status match
        case 1 => throw IllegalArgumentException(msg)          // `require` failures (Mat.scala:2816-2817 ...)
        case 2 => throw ArithmeticException(msg)               // singular matrix (Mat.scala:990)
        case 3 => throw UnsupportedOperationException(msg)     // empty min/max (Mat.scala:2216)
        case 6 => throw OutOfMemoryError(msg)
        case _ => throw RuntimeException(s"libunimat status $status: $msg")
