#!/usr/bin/env python3
"""bench.py -- 3PRF closed-form fit on the BASELINE.json config-4 panel (T=8192 x N=262144 x L=8,
FP64, synthetic, 16 GiB), column-sharded over --gpus N B200s; plus (N=1) the config-2 MatD
elementwise / reduction ops and a config-3 matmul sweep as extras.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference algorithm on the host cores

Prints ONE JSON line (rank 0).  `value` = algorithmic FP64 TFLOP/s of the whole job
(T*N*(4L+6) flop per fit) with inputs resident in HBM; `ms_per_step` is the
fit time; `e2e` is the same fit through the C-ABI host entry point (`um_tprf_fit_host`: pinned
host panel streamed in, forecasts/alpha/R^2 copied out, all inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_ROWS, N_COLS, L_PROX = 8192, 262144, 8
METRIC = "3prf_closed_form_fit_fp64_tflops"
UNIT = "TFLOP/s"


FP64_PEAK = 37.0   # TFLOP/s: DMMA m8n8k4 / DFMA issue peak measured on this pool (tools/microbench/fp64_peak.cu, profiles/r1_fp64_peak.log)


def tprf_flops(T, N, L):
    return float(T) * float(N) * (4.0 * L + 6.0)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU legs: the reference's algorithm on the host cores.  No JVM on the box, so the BLAS path is the
# NumPy/OpenBLAS restatement in oracle/ (pinned to the reference's fixtures) and the pure-JVM path is the C++
# restatement oracle/jvm_restated.cpp.  These are reported baselines, never on the product path.
# ------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


WORKLOAD = "3PRF closed-form fit, T=8192 x N=262144 x L=8 FP64 panel (BASELINE config 4), sharded by predictor column"


def bench_config(T, N, Lp, world):
    """The `config` object of BOTH arms (the driver compares them)."""
    return {"workload": WORKLOAD, "T": T, "N": N, "L": Lp, "cols_per_gpu": N // world,
            "l2": "inputs (16 GiB / n_gpus per rank) are larger than L2; no flush needed",
            "parallelism": f"column-shard x{world}, one packed all-reduce per fit"}


def cpu_shard_pass(T, N, Lp, shard_cols, budget_s, seed=0):
    """One pass of the reference's closed form over the panel, AS COLUMN SHARDS SUMMED (the same decomposition the GPUs
    use: oracle.closed_form_partials per shard, one sum, closed_form_from_partials), for as many shards of `shard_cols`
    columns as fit in `budget_s` seconds of host time.  Shard k is columns [k*shard_cols, (k+1)*shard_cols) of the
    seeded panel, regenerated per shard (outside the timed part) so no 16 GiB host array is needed.
    Returns (seconds of fit work, columns processed, R^2 of the processed columns)."""
    import numpy as np
    from oracle import tprf_oracle as to
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, Lp))
    total, spent, done = None, 0.0, 0
    while done < N:
        w = min(shard_cols, N - done)
        X = np.random.default_rng([seed, done]).standard_normal((T, w))
        t0 = time.perf_counter()
        part = to.closed_form_partials(X, Z, T)
        total = part if total is None else {k: total[k] + part[k] for k in part if k != "W0"}
        spent += time.perf_counter() - t0
        done += w
        if spent >= budget_s:
            break
    t0 = time.perf_counter()
    yhat = to.closed_form_from_partials(y, total, done)[0]
    r2 = 1.0 - float(((y - yhat) ** 2).sum()) / float(((y - y.mean()) ** 2).sum())
    spent += time.perf_counter() - t0
    return spent, done, r2


def run_reference(args, rank, world):
    """`--impl reference`: the reference's closed form on the host cores, on OUR arm's config / metric / unit.  A step is
    one bounded sample of the workload -- `REF_SHARD` columns of the panel per step, a different shard each step, through
    the column-shard decomposition (partials summed) -- so `ms_per_step` is MEASURED, never extrapolated, and the whole
    --steps/--warmup run ends within a few minutes; `value` (TFLOP/s) is a rate and does not depend on the sample size."""
    if rank != 0:
        return
    import numpy as np
    from oracle import tprf_oracle as to
    T, N, Lp = T_ROWS, N_COLS, L_PROX
    cores = host_cores()
    REF_SHARD = 4096
    rng = np.random.default_rng(0)
    y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, Lp))
    shard = lambda k: np.random.default_rng([0, k * REF_SHARD]).standard_normal((T, REF_SHARD))
    X = shard(0)
    for _ in range(max(1, args.warmup)):
        to.closed_form_partials(X, Z, T)
    total, dt = None, 0.0
    for k in range(args.steps):
        X = shard(k % (N // REF_SHARD))
        t0 = time.perf_counter()
        part = to.closed_form_partials(X, Z, T)
        total = part if total is None else {q: total[q] + part[q] for q in part if q != "W0"}
        dt += time.perf_counter() - t0
    t0 = time.perf_counter()
    yhat = to.closed_form_from_partials(y, total, args.steps * REF_SHARD)[0]
    r2 = 1.0 - float(((y - yhat) ** 2).sum()) / float(((y - y.mean()) ** 2).sum())
    dt += time.perf_counter() - t0
    dt /= args.steps
    val = tprf_flops(T, REF_SHARD, Lp) / dt / 1e12
    sample = (f"closed form as column shards summed (oracle/tprf_oracle.py closed_form_partials -> closed_form_from_partials, "
              f"NumPy+OpenBLAS, {cores} host threads): one {REF_SHARD}-column shard of the {N}-column panel per step, "
              f"{args.steps} different shards; at this rate the whole panel takes {dt * (N / REF_SHARD):.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(T, N, Lp, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "r_squared_of_sampled_columns": r2},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def best_of(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_matd_ops(n=8192):
    """CPU figures beside BASELINE config 2, on a bounded n x n sample of the 32768^2 workload (GB/s is a rate): the BLAS /
    NumPy path (what the reference's own py/bench.py times, py/bench.py:37-127) and the restated pure-JVM kernels."""
    import numpy as np
    from oracle import jvm_restated as jr
    cores = host_cores()
    E = float(n) * n
    m = np.random.default_rng(42).standard_normal((n, n))
    out = np.empty_like(m)
    rowv = m.mean(0)
    res = {}

    def put(name, alg_bytes, np_fn, jr_fn=None, jr_threads=None):
        e = {}
        t = best_of(np_fn)
        e["numpy"] = {"ms": round(t * 1e3, 2), "gbs": round(alg_bytes / t / 1e9, 2)}
        if jr_fn is not None:
            t = best_of(jr_fn)
            e["jvm_restated"] = {"ms": round(t * 1e3, 2), "gbs": round(alg_bytes / t / 1e9, 2),
                                 "threads": cores if jr_threads is None else jr_threads}
        res[name] = e

    put("sub_rowvec(m - m.mean(0))", 16 * E, lambda: np.subtract(m, rowv, out=out), lambda: jr.sub_rowvec(m, rowv, out, cores))
    put("add_same_shape", 24 * E, lambda: np.add(m, m, out=out), lambda: jr.add_same(m, m, out, cores))
    put("sigmoid", 16 * E, lambda: np.divide(1.0, 1.0 + np.exp(-m), out=out), lambda: jr.sigmoid(m, out, cores))

    def np_softmax(ax):
        e = np.exp(m - m.max(axis=ax, keepdims=True))
        return e / e.sum(axis=ax, keepdims=True)
    put("softmax_axis1", 16 * E, lambda: np_softmax(1), lambda: jr.softmax_rows(m, out), 1)
    put("softmax_axis0", 16 * E, lambda: np_softmax(0))
    put("sum_axis0", 8 * E, lambda: m.sum(0), lambda: jr.col_sums(m, cores))
    put("sum_axis1", 8 * E, lambda: m.sum(1), lambda: jr.row_sums(m, cores))
    put("mean_axis0", 8 * E, lambda: m.mean(0))
    put("var_axis0", 8 * E, lambda: m.var(0))
    put("var_axis1", 8 * E, lambda: m.var(1))
    put("std_axis0", 8 * E, lambda: m.std(0))
    put("sum_all", 8 * E, lambda: m.sum(), lambda: jr.total(m, cores))
    put("argmax_all", 8 * E, lambda: m.argmax())
    put("gt_mask", 9 * E, lambda: m > 0.0)
    put("cumsum_axis0", 16 * E, lambda: np.cumsum(m, axis=0, out=out))
    return {"cores": cores, "sample": f"{n} x {n} FP64 (a 1/{(32768 // n) ** 2} sample of the 32768^2 matrix; GB/s against the same "
                                      "algorithmic bytes as the GPU rows); numpy = NumPy + OpenBLAS default threads, jvm_restated = "
                                      "oracle/jvm_restated.cpp (g++ -O2 -ffp-contract=off, the reference's ForkJoin chunking)",
            "ops": res}


def cpu_matmul(sizes=(1024, 2048, 4096), pure_sizes=(1024,)):
    import numpy as np
    from oracle import jvm_restated as jr
    cores = host_cores()
    res = {}
    for n in sizes:
        a = np.random.default_rng(42).standard_normal((n, n)); b = np.random.default_rng(43).standard_normal((n, n))
        t = best_of(lambda: a @ b, reps=2)
        res[f"f64_{n}"] = {"numpy_openblas": {"ms": round(t * 1e3, 2), "tflops": round(2.0 * n ** 3 / t / 1e12, 4)}}
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        t = best_of(lambda: a32 @ b32, reps=2)
        res[f"f32_{n}"] = {"numpy_openblas": {"ms": round(t * 1e3, 2), "tflops": round(2.0 * n ** 3 / t / 1e12, 4)}}
        if n in pure_sizes:
            t = best_of(lambda: jr.matmul_pure(a, b, cores), reps=2)
            res[f"f64_{n}"]["jvm_restated_matmulPure"] = {"ms": round(t * 1e3, 2), "tflops": round(2.0 * n ** 3 / t / 1e12, 4), "threads": cores}
    return {"cores": cores, "note": "CPU rows stop at 4096 (a 16384^3 FP64 product is ~20 s of OpenBLAS): larger GPU rows are set beside the "
                                    "4096 rate; matmulPure = MatmulPure.multiply restated (pack4 + 4x4 microkernel)", "sizes": res}


def cublas_dgemm_tflops(u, sizes):
    """cuBLAS DGEMM on the same box: a YARDSTICK for the FP64 `*@` rows, not a code path of the library."""
    import ctypes as C_
    try:
        cb = C_.CDLL("libcublas.so.12"); rt = C_.CDLL("libcudart.so.12")
    except OSError as ex:
        return {"error": str(ex)[:120]}
    h = C_.c_void_p()
    if cb.cublasCreate_v2(C_.byref(h)) != 0:
        return {"error": "cublasCreate failed"}
    out = {}
    one, zero = C_.c_double(1.0), C_.c_double(0.0)
    for n in sizes:
        a = u.MatD.randn(n, n, seed=42); b = u.MatD.randn(n, n, seed=43); c = u.MatD.zeros(n, n)
        u.sync()
        call = lambda: cb.cublasDgemm_v2(h, 0, 0, n, n, n, C_.byref(one), C_.c_void_p(b._buf.ptr), n, C_.c_void_p(a._buf.ptr), n,
                                         C_.byref(zero), C_.c_void_p(c._buf.ptr), n)   # row-major C = A B  <=>  column-major C' = B' A'
        reps = 3 if n >= 8192 else 10
        for _ in range(2):
            call()
        rt.cudaDeviceSynchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        rt.cudaDeviceSynchronize()
        t = (time.perf_counter() - t0) / reps
        out[str(n)] = {"ms": round(t * 1e3, 3), "tflops": round(2.0 * n ** 3 / t / 1e12, 2)}
        del a, b, c
    cb.cublasDestroy_v2(h)
    return out


def oos_rows(u, small=False):
    """estimate3prf's out-of-sample procedures, every window of a procedure in one launch
    (um_tprf_oos_windows).  650 x 40 x 2 is the reference's published shape (README.md:315-317: Scala 1.65 ms Recursive /
    5.7 ms Cross Val); the wide panel is this repo's.  GPU figure = whole host call (window setup, launch, copy-back, R^2),
    best of 7 by wall clock; cpu = the NumPy restatement (oracle.estimate3prf, one run) on the same input."""
    import numpy as np
    from oracle import tprf_oracle as to
    out = {}
    shapes = ((650, 40, 2, True),) if small else ((650, 40, 2, True), (200, 16384, 4, True))   # the wide CPU runs take ~2 + 7 s
    for T, N, Lp, with_cpu in shapes:
        rng = np.random.default_rng(1)
        X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, Lp))
        dy, dX, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)
        for proc in ("OOS Recursive", "OOS Cross Val"):
            call = lambda: u.estimate3prf(dy, dX, dZ, proc)
            r = call(); u.sync()
            best = 1e30
            for _ in range(7):
                t0 = time.perf_counter(); r = call(); u.sync(); best = min(best, time.perf_counter() - t0)
            nwin = int(np.isfinite(r.forecasts.to_numpy()).sum())
            e = {"ms": round(best * 1e3, 3), "windows": nwin, "windows_per_s": round(nwin / best, 1), "r_squared": r.rSquared}
            if with_cpu:
                t0 = time.perf_counter(); o = to.estimate3prf(y, X, Z, proc); t = time.perf_counter() - t0
                e["cpu"] = {"numpy_port_ms": round(t * 1e3, 1), "cores": host_cores(),
                            "max_abs_forecast_diff": float(np.nanmax(np.abs(o.forecasts - r.forecasts.to_numpy())))}
            out[f"{proc} T={T} N={N} L={Lp}"] = e
    out["note"] = ("recorded on the build container (profiles/r2_reference_python_cpu.json, 8 cores): the unmodified reference "
                   "py/tprf3fast.py takes 75 ms (Recursive) / 102 ms (Cross Val) at 650 x 40 x 2; the reference README publishes "
                   "1.65 / 5.7 ms for its Scala implementation")
    return out


def cpu_config5(panels=16):
    import numpy as np
    from oracle import tprf_oracle as to
    T, N, Lp = 512, 2048, 4
    data = []
    for p in range(panels):
        rng = np.random.default_rng(p)
        data.append((rng.standard_normal((T, N)), rng.standard_normal((T, 1)), rng.standard_normal((T, Lp))))
    to.tprf_closed_form_gram(data[0][1], data[0][0], data[0][2])
    t0 = time.perf_counter()
    for X, y, Z in data:
        to.tprf_closed_form_gram(y, X, Z)
    t = time.perf_counter() - t0
    return {"cores": host_cores(), "panels_per_s": round(panels / t, 1), "gbs": round(8.0 * T * N * panels / t / 1e9, 2),
            "sample": f"{panels} of the 4096 panels, oracle Gram closed form (NumPy+OpenBLAS) one panel after the other"}


# ------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------
def time_op(u, fn, warm=16, reps=60):
    """min over `reps` CUDA-event timings after `warm` untimed calls: SURVEY.md 8d's protocol for config 2 (warm-up 16,
    min-of-60, mirroring the reference's py/bench.py:40-55)."""
    from uni_b200._lib import lib, check
    for _ in range(warm):
        fn()
    best = 1e30
    ms = C.c_float(0)
    for _ in range(reps):
        check(lib.um_timer_start())
        fn()
        check(lib.um_timer_stop(C.byref(ms)))
        best = min(best, ms.value)
    return best


def numpy_stream_panel(u, T, N, Lp, c0, c1, seed=0):
    """The reference benches' panel -- numpy.random.default_rng(seed).standard_normal, drawn X (T x N), then y (T x 1),
    then Z (T x L) (apps/Tprf3Bench.scala:146-149, py/bench_tprf3.py:46-49) -- generated in HBM by the device PCG64 /
    ziggurat generator (csrc/rng.cu: the same stream NumPy produces), in row chunks of at most 1 GiB; a rank keeps the
    columns [c0, c1) of every row, so all ranks see shards of ONE panel without any host array."""
    from uni_b200 import _lib as L_
    from uni_b200._lib import check, lib
    g = u.NumPyRNG(seed)
    Nl = c1 - c0
    X = u.Mat._alloc(T, Nl, L_.F64)
    rows_per = max(1, (1 << 27) // N)
    for r0 in range(0, T, rows_per):
        r1 = min(T, r0 + rows_per)
        chunk = g.randn(r1 - r0, N)
        check(lib.um_copy(chunk._raw_slice(0, r1 - r0, c0, c1)._ref(), X._raw_slice(r0, r1, 0, Nl)._ref()))
        del chunk
    y = g.randn(T, 1)
    Z = g.randn(T, Lp)
    return X, y, Z


def matd_ops(u, peak_gbs, n=32768):
    """BASELINE config 2: ops on a 32768^2 FP64 matrix, warm-up 16 / min-of-60 CUDA-event timings; GB/s against the
    ALGORITHMIC bytes of SURVEY.md section 8d (E = n^2 elements).  Every op's inputs and outputs (8 GiB each) are larger
    than L2, so no flush is needed between repetitions."""
    E = float(n) * n
    m = u.MatD.randn(n, n, seed=42)
    rowv = m.mean(0)
    colv = m.mean(1)
    out = {}

    def rec(name, fn, alg_bytes, note=None):
        time.sleep(1.0)   # every op is measured from the same rested state (76 back-to-back 8 GiB ops heat the part)
        ms = time_op(u, fn)
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": round(ms, 4), "gbs": round(gbs, 1), "frac_hbm": round(gbs / peak_gbs, 4)}
        if note:
            out[name]["note"] = note

    rec("sub_rowvec(m - m.mean(0))", lambda: m - rowv, 16 * E)
    rec("sub_colvec(m - m.mean(1))", lambda: m - colv, 16 * E)
    m2 = u.MatD.randn(n, n, seed=43)
    rec("add_same_shape", lambda: m + m2, 24 * E)
    del m2
    rec("mul_scalar", lambda: m * 1.5, 16 * E)
    rec("neg", lambda: -m, 16 * E)
    rec("relu", lambda: m.relu(), 16 * E)
    rec("sigmoid", lambda: m.sigmoid(), 16 * E)
    rec("softmax_axis1", lambda: m.softmax(1), 16 * E, "single read: the 256 KB row held in registers + shared memory across a 2-CTA cluster")
    rec("softmax_axis0", lambda: m.softmax(0), 16 * E,
        "single read for 85 % of the columns: 16-column strips held in tensor memory across 16-CTA clusters (softmax_strip.cuh); "
        "the other 15 % run through the two-read kernels at the same time on the SMs and issue slots the clusters leave idle")
    rec("sum_axis0", lambda: m.sum(0), 8 * E)
    rec("sum_axis1", lambda: m.sum(1), 8 * E)
    rec("mean_axis0", lambda: m.mean(0), 8 * E)
    rec("mean_axis1", lambda: m.mean(1), 8 * E)
    rec("var_axis0", lambda: m.var(0), 8 * E, "single read (shifted sums + Chan merge)")
    rec("var_axis1", lambda: m.var(1), 8 * E, "single read (shifted sums + Chan merge)")
    rec("std_axis0", lambda: m.std(0), 8 * E)
    rec("sum_all", lambda: m.sum(), 8 * E)
    rec("argmax_all", lambda: m.argmax(), 8 * E)
    rec("gt_mask", lambda: m.gt(0.0), 9 * E)
    # SURVEY 8f rank 2 (running folds are sequential per lane, as the reference pins them; parallel across lanes)
    rec("cumsum_axis0", lambda: m.cumsum(0), 16 * E)
    rec("cumsum_axis1", lambda: m.cumsum(1), 16 * E)
    rec("cummax_axis0", lambda: m.cummax(0), 16 * E)
    rec("cummax_axis1", lambda: m.cummax(1), 16 * E)
    mask = m.gt(0.0)
    rec("where(mask, m, -m)", lambda: u.Mat.where(mask, m, m), 17 * E, "x and y alias the same matrix: one 8-byte read per element")
    rec("select m(mask)", lambda: m[mask], 14 * E, "mask read twice (count, scatter) + operand + ~half the elements written; includes the host read of k")
    del mask
    # the NumPy-compatible PCG64 stream generated in HBM (8E bytes written; randn regenerates the
    # stream twice -- tabulate, then emit -- and synchronises to learn how many draws it consumed)
    g = u.NumPyRNG(0)
    rec("rng_rand(PCG64)", lambda: g.rand(n, n), 8 * E)
    rec("rng_randn(PCG64 ziggurat)", lambda: g.randn(n, n), 8 * E, "bit-identical to numpy default_rng.standard_normal; three kernels")
    t = m.T
    rec("sum_axis0@transposed_view", lambda: t.sum(0), 8 * E)
    rec("sub_rowvec@row_slice_view", lambda: m.slice(1024, n, 0, n) - rowv, 16 * (E - 1024.0 * n))

    def inplace():
        m.__iadd__(0.0)
    rec("inplace_add_scalar(:+=)", inplace, 16 * E)
    return out


def matmul_sweep(u, sizes=(1024, 2048, 4096, 8192)):
    out = {}
    for dt_name, fac in (("f64", u.MatD), ("f32", u.MatF)):
        for n in sizes:
            a = fac.randn(n, n, seed=42); b = fac.randn(n, n, seed=43)
            big = n >= 8192
            ms = time_op(u, lambda: a @ b, warm=2 if big else 5, reps=3 if big else 15)
            tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
            e = {"ms": round(ms, 3), "tflops": round(tf, 2)}
            if dt_name == "f64":
                e["frac_fp64_peak"] = round(tf / FP64_PEAK, 4)          # tensor-pipe roofline: measured DMMA issue peak
            out[f"{dt_name}_{n}"] = e
            if dt_name == "f64" and n <= 4096:
                at = a.T
                ms = time_op(u, lambda: at @ b, warm=5, reps=15)
                tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
                out[f"{dt_name}_{n}_transposed_left"] = {"ms": round(ms, 3), "tflops": round(tf, 2), "frac_fp64_peak": round(tf / FP64_PEAK, 4)}
            del a, b
    return out


def batched_config5(u, rank, world, max_over_ranks, barrier, peak_gbs, small=False):
    """BASELINE config 5: 4096 independent panels (T=512, N=2048, L=4), panels sharded across ranks, no collective."""
    from uni_b200 import _lib as L_
    from uni_b200._lib import check, lib
    T, N, Lp, total = 512, 2048, 4, (256 if small else 4096)
    nb = total // world + (1 if rank < total % world else 0)
    X = u.MatD.randn(nb * T, N, seed=7000 + rank)
    y = u.MatD.randn(nb * T, 1, seed=7100 + rank)
    Z = u.MatD.randn(nb * T, Lp, seed=7200 + rank)
    f = u.MatD.zeros(nb * T, 1); r2 = u.MatD.zeros(nb, 1)
    p = lambda m: C.c_void_p(m._buf.ptr)
    call = lambda: check(lib.um_tprf_fit_batched(L_.TPRF_CLOSED, p(y), p(X), p(Z), T, N, Lp, nb, p(f), None, p(r2)))
    for _ in range(3):
        call()
    barrier()
    ms = C.c_float(0)
    reps = 5
    check(lib.um_timer_start())
    for _ in range(reps):
        call()
    check(lib.um_timer_stop(C.byref(ms)))
    barrier()
    t = max_over_ranks(ms.value / reps)
    gbs = 8.0 * T * N * nb / (t * 1e-3) / 1e9
    return {"panels": total, "panels_per_gpu": nb, "ms": round(t, 4), "panels_per_s": round(total / (t * 1e-3), 1),
            "gbs_per_gpu": round(gbs, 1), "frac_hbm": round(gbs / peak_gbs, 4),
            "tflops": round(total * tprf_flops(T, N, Lp) / (t * 1e-3) / 1e12, 2)}


def run_ours(args, rank, world, local_rank):
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    import numpy as np
    import uni_b200 as u
    from uni_b200 import _lib as L_
    from uni_b200._lib import check, lib
    from uni_b200.dist import init_comm, shard_range

    if u.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (uni_b200 has no CPU fallback)")
    u.init(local_rank)
    T, N, Lp = T_ROWS, N_COLS, L_PROX
    if args.small:
        T, N = 2048, 16384
    # N > 1: rank 0 first fits the WHOLE panel on its one GPU (no communicator yet), so the line can carry the distance
    # between the sharded fit and the single-GPU fit of the same panel (parity at full size, visible to the driver)
    one_gpu = None
    if world > 1 and rank == 0 and not args.debug_no_comm:
        Xf, yf, Zf = numpy_stream_panel(u, T, N, Lp, 0, N, seed=0)
        r1 = u.tprf3._fit(L_.TPRF_CLOSED, yf, Xf, Zf, n_total=N, want_all=False)
        one_gpu = (r1.forecasts.to_numpy().ravel().copy(), r1.rSquared)
        del Xf, yf, Zf, r1
    if world > 1:
        def bcast(b):
            import torch
            t = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, src=0)
            return bytes(t.cpu().tolist())
        if not args.debug_no_comm:
            init_comm(rank, world, bcast)

    def barrier():
        u.sync()
        if dist is not None:
            import torch
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c0, c1 = shard_range(N, rank, world)
    Nl = c1 - c0
    try:
        X, y, Z = numpy_stream_panel(u, T, N, Lp, c0, c1, seed=0)
        data_note = ("synthetic: numpy.random.default_rng(0).standard_normal drawn X, y, Z (apps/Tprf3Bench.scala:146-149), "
                     "generated in HBM by the device PCG64/ziggurat generator; each rank keeps its column shard")
    except Exception as ex:  # the counter-based generator needs nothing but one launch
        X = u.MatD.randn(T, Nl, seed=1000 + rank)
        y = u.MatD.randn(T, 1, seed=2)
        Z = u.MatD.randn(T, Lp, seed=3)
        data_note = f"synthetic: counter-based N(0,1) (numpy-stream generation failed: {str(ex)[:80]})"
    fit = lambda: u.tprf3._fit(L_.TPRF_CLOSED, y, X, Z, n_total=N, want_all=False)
    peak_gbs, peak_src = load_peaks()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # samples from here on: warm-up and timed steps are the same workload, back to back
    warm_steps = max(args.warmup, 3)
    for _ in range(warm_steps):
        r = fit()
    barrier()
    n_launch0 = u.launch_count()
    ms = C.c_float(0)
    t0 = time.perf_counter()
    check(lib.um_timer_start())
    for _ in range(args.steps):
        r = fit()
    check(lib.um_timer_stop(C.byref(ms)))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = u.launch_count() - n_launch0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms.value / args.steps)
    flops = tprf_flops(T, N, Lp)
    value = flops / (ms_step * 1e-3) / 1e12

    # ---- SURVEY 8d / apps/Tprf3Bench.scala:186-203: median of 25 individually timed fits (the contract's ms_per_step above is
    # the mean over --steps back-to-back fits; both are reported)
    # (the sweep draws enough power that a B200 sheds ~7 % of its speed after about a second of back-to-back fits --
    # tools/iter_vs_closed.py: 3.87 -> 4.13 ms; every figure of this section starts from the same rested state)
    time.sleep(3.0)
    per_fit = []
    for _ in range(25):
        check(lib.um_timer_start())
        fit()
        check(lib.um_timer_stop(C.byref(ms)))
        per_fit.append(max_over_ranks(ms.value))
    per_fit.sort()
    median_ms = per_fit[len(per_fit) // 2]

    # ---- the three-pass (iterative, Tprf3.t3prf) fit of the same panel: BASELINE config 4 names both fits
    fit_iter = lambda: u.tprf3._fit(L_.TPRF_ITER, y, X, Z, n_total=N, want_all=False)
    time.sleep(3.0)
    for _ in range(2):
        fit_iter()
    barrier()
    check(lib.um_timer_start())
    for _ in range(args.steps):
        fit_iter()
    check(lib.um_timer_stop(C.byref(ms)))
    barrier()
    it_ms = max_over_ranks(ms.value / args.steps)
    iterative = {"ms_per_fit": round(it_ms, 4), "tflops": round(flops / (it_ms * 1e-3) / 1e12, 3), "steps": args.steps,
                 "note": "Tprf3.t3prf on the same resident panel: same single-read sweep, three-pass finish"}

    # ---- more proxies than one DMMA fragment holds (the reference's fits take any L): one fused sweep per group of 8
    # proxies + operator finish (csrc/tprf_wide.cu), here L = 16 on the same resident panel
    try:
        Z16 = u.MatD.randn(T, 16, seed=77)
        fit16 = lambda: u.tprf3._fit(L_.TPRF_CLOSED, y, X, Z16, n_total=N, want_all=False)
        time.sleep(1.0)
        fit16()
        barrier()
        check(lib.um_timer_start())
        for _ in range(3):
            r16 = fit16()
        check(lib.um_timer_stop(C.byref(ms)))
        barrier()
        w_ms = max_over_ranks(ms.value / 3)
        wide = {"L": 16, "ms_per_fit": round(w_ms, 4), "tflops": round(tprf_flops(T, N, 16) / (w_ms * 1e-3) / 1e12, 3),
                "x_reads": 2, "r_squared": r16.rSquared}
        del Z16, r16
    except Exception as ex:
        wide = {"error": str(ex)[:200]}

    # ---- per-kernel device times (outside the timed region) -> roofline of the dominant kernel
    check(lib.um_prof_enable(1)); check(lib.um_prof_reset())
    PROF_N = 5
    for _ in range(PROF_N):
        fit()
    names = ["colstats", "colfinal", "project", "gather", "finish", "outputs", "allreduce", "prep"]
    kt = {}
    for i, nm in enumerate(names):
        tot, cnt = C.c_double(0), C.c_int64(0)
        check(lib.um_prof_read(i, C.byref(tot), C.byref(cnt)))
        if cnt.value:
            kt[nm] = tot.value / cnt.value
    check(lib.um_prof_enable(0))
    # slot 0 ("colstats") is the single-read cluster sweep when no "project" sweep ran
    fused = "project" not in kt
    if fused:
        kt["fused_sweep"] = kt.pop("colstats")
        dom = "fused_sweep"
    else:
        dom = max(("colstats", "project"), key=lambda k: kt.get(k, 0.0))
    dom_ms = max_over_ranks(kt[dom])
    alg_bytes = 8.0 * T * Nl
    # The sweep is two kernels running CONCURRENTLY (clusters + the grid sweep on the SMs the cluster shape strands), so
    # "the kernel's launch duration" is taken from the timed region itself: the sweep's share of the profiled fit
    # (events around the pair, both kernels inside) applied to ms_per_step -- never more than ms_per_step.
    prof_fit_ms = sum(v for k, v in kt.items() if k != "outputs")
    sweep_ms = ms_step * min(1.0, dom_ms / prof_fit_ms) if prof_fit_ms > 0 else ms_step
    achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes of the same kernel and shape from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_tprf_fused_traffic.json")))
        if fused and tr["T"] == T and tr["N_local"] == Nl:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    k_tflops = tprf_flops(T, Nl, Lp) / (sweep_ms * 1e-3) / 1e12
    g_, r_, n_ = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib.um_tprf_fused_info(T, C.byref(g_), C.byref(r_), C.byref(n_)))
    roofline = {
        "bound": "hbm", "kernel": "k_tprf_fused" if fused else f"k_tprf_{dom}", "achieved": round(achieved, 1), "peak": peak_gbs,
        "unit": "GB/s", "frac": round(achieved / peak_gbs, 4), "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes,
        "fp64": {"achieved": round(k_tflops, 2), "peak": FP64_PEAK, "unit": "TFLOP/s", "frac": round(k_tflops / FP64_PEAK, 4),
                 "peak_source": "measured DMMA m8n8k4 / DFMA issue peak (tools/microbench/fp64_peak.cu, profiles/r1_fp64_peak.log); "
                                "the kernel issues 20 FP64 slots per element, counted as 4L+6 = 38 flop"},
        "sweep_ms": round(sweep_ms, 4), "sweep_ms_note": "ms_per_step x the sweep's share of the event-timed fit; <= ms_per_step",
        "kernel_ms": {k: round(v, 4) for k, v in kt.items()},
        "fit_frac_x_read_once": round((8.0 * T * N / world) / (ms_step * 1e-3) / 1e9 / peak_gbs, 4),
        "sweep": ({"path": "single-read cluster sweep", "cluster_ctas": g_.value, "rows_per_cta": r_.value,
                   "resident_clusters": n_.value, "sms_used": g_.value * n_.value} if fused else {"path": "two sweeps"}),
        "note": ("X is read from HBM once (8*T*N_local bytes per launch); the sweep is bound by FP64 issue and HBM almost "
                 "equally (DESIGN.md 4.2)" if fused else
                 "each of the two sweeps (colstats, project) reads the shard once: 8*T*N_local bytes per launch; "
                 "fit_frac_x_read_once charges the whole fit with ONE read of X (SURVEY 8d)"),
    }

    # ---- end to end through the C-ABI host entry point (pinned host panel, H2D inside the timed region)
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped by --no-e2e (profiling run)")
        nbytes = T * Nl * 8
        hx = C.c_void_p()
        check(lib.um_malloc_host(nbytes, C.byref(hx)))
        check(lib.um_d2h(hx, C.c_void_p(X._buf.ptr), nbytes))
        yh = y.to_numpy(); zh = Z.to_numpy()
        fh = np.zeros(T); ah = np.zeros(Nl); r2 = C.c_double(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        call = lambda: check(lib.um_tprf_fit_host(L_.TPRF_CLOSED, vp(yh), hx, Nl, vp(zh), T, Nl, Lp, N, vp(fh), vp(ah), C.byref(r2)))
        call()
        barrier()
        esteps = max(1, min(args.steps, 3))
        t1 = time.perf_counter()
        for _ in range(esteps):
            call()
        barrier()
        e_ms = max_over_ranks((time.perf_counter() - t1) * 1e3 / esteps)
        dev_f = r.forecasts.to_numpy().ravel()
        e2e = {"value": flops / (e_ms * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": e_ms, "steps": esteps,
               "h2d_bytes_per_step": nbytes + T * (Lp + 1) * 8, "d2h_bytes_per_step": T * 8 + Nl * 8 + 16,
               "api": "um_tprf_fit_host (C ABI, host pointers)",
               "max_abs_diff_vs_device_resident_fit": float(np.max(np.abs(dev_f - fh))) if world == 1 else None}
        check(lib.um_free_host(hx))
    except Exception as ex:  # e.g. the box cannot pin 16 GiB
        e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    extras = {"config4_iterative_t3prf": iterative, "config4_closed_form_16_proxies": wide}
    if not args.skip_ops:
        try:
            extras["batched_config5"] = batched_config5(u, rank, world, max_over_ranks, barrier, peak_gbs, small=args.small)
        except Exception as ex:
            extras["batched_config5"] = {"error": str(ex)[:200]}
    cpu_baseline = None
    if rank == 0 and world == 1:
        # the reference's closed form on this box's host cores: the same workload through the column-shard decomposition,
        # as many 4096-column shards of the panel as fit in ~15 s (measured, not extrapolated; the rate is what is reported)
        cores = host_cores()
        dt, ncols, r2cpu = cpu_shard_pass(T, N, Lp, 4096, 4.0 if args.small else 15.0)
        cpu_baseline = {"value": tprf_flops(T, ncols, Lp) / dt / 1e12, "unit": UNIT, "cores": cores, "kind": "port",
                        "seconds": round(dt, 2), "columns": ncols,
                        "whole_panel_seconds_at_this_rate": round(dt * N / ncols, 1),
                        "sample": (f"closed form as column shards summed (oracle closed_form_partials / closed_form_from_partials, NumPy+OpenBLAS, "
                                   f"{cores} threads) over the first {ncols} of {N} columns of a seeded panel, T={T}, L={Lp}")}
        if not args.skip_ops:
            del X
            try:
                ops = matd_ops(u, peak_gbs, 4096 if args.small else 32768)
                cpu_ops = cpu_matd_ops(2048 if args.small else 8192)
                for k, v in ops.items():          # every GPU row carries the CPU figure of the same op where one exists
                    if k in cpu_ops["ops"]:
                        v["cpu"] = dict(cpu_ops["ops"][k], cores=cpu_ops["cores"])
                extras["matd_ops_32768x32768_f64"] = ops
                extras["matd_ops_cpu_note"] = cpu_ops["sample"]
                sizes = (1024, 2048) if args.small else (1024, 2048, 4096, 8192, 16384)
                mm = matmul_sweep(u, sizes)
                yard = cublas_dgemm_tflops(u, sizes)
                cpu_mm = cpu_matmul((1024, 2048) if args.small else (1024, 2048, 4096))
                for k, v in mm.items():
                    parts = k.split("_")
                    n_ = parts[1]
                    if parts[0] == "f64" and isinstance(yard.get(n_), dict):
                        v["cublas_dgemm_tflops"] = yard[n_]["tflops"]
                        v["vs_cublas"] = round(v["tflops"] / yard[n_]["tflops"], 3)
                    ck = f"{parts[0]}_{n_}" if f"{parts[0]}_{n_}" in cpu_mm["sizes"] else f"{parts[0]}_{max(int(q.split('_')[1]) for q in cpu_mm['sizes'])}"
                    v["cpu"] = dict(cpu_mm["sizes"][ck], cores=cpu_mm["cores"], at_n=int(ck.split("_")[1]))
                extras["matmul_sweep"] = mm
                extras["matmul_sweep_note"] = ("frac_fp64_peak = TFLOP/s / 37 (measured DMMA issue peak); cublas_dgemm_tflops = cuBLAS on the "
                                               "same box, a yardstick only; " + cpu_mm["note"])
                if isinstance(extras.get("batched_config5"), dict) and "error" not in extras["batched_config5"]:
                    extras["batched_config5"]["cpu"] = cpu_config5()
                extras["oos_procedures"] = oos_rows(u, small=args.small)
            except Exception as ex:
                extras["ops_error"] = str(ex)[:300]
    exchange = "none (1 GPU)"
    if world > 1 and not args.debug_no_comm:
        p2p_ = C.c_int32(0)
        check(lib.um_comm_transport(C.byref(p2p_)))
        exchange = "peer memory (CUDA IPC over NVLink), one kernel" if p2p_.value else "NCCL all-gather + rank-order sum"
    parity = None
    if rank == 0:
        import hashlib
        fN = r.forecasts.to_numpy().ravel()
        parity = {"forecast_sha256_16": hashlib.sha256(fN.tobytes()).hexdigest()[:16], "r_squared": r.rSquared}
        if one_gpu is not None:   # the sharded fit against the single-GPU fit of the same panel, at full size
            parity["max_abs_diff_vs_1gpu"] = float(np.max(np.abs(fN - one_gpu[0])))
            parity["r_squared_1gpu"] = one_gpu[1]
            parity["r_squared_abs_diff_vs_1gpu"] = abs(r.rSquared - one_gpu[1])
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm_steps,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": data_note,
            "config": bench_config(T, N, Lp, world), "exchange": exchange,
            "ms_per_fit_median_of_25": median_ms, "parity": parity,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "wall_ms_per_step": wall_ms / args.steps, "r_squared": r.rSquared,
        }
        line.update(extras)
        if args.debug_no_comm:
            line["DIAGNOSTIC"] = "--debug-no-comm: shards fitted without the all-reduce; not a benchmark value"
        print(json.dumps(line), flush=True)
    if dist is not None:
        from uni_b200.dist import destroy_comm
        barrier()
        if not args.debug_no_comm:
            destroy_comm()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-ops", action="store_true", help="skip the config-2/3 extras (MatD ops, matmul sweep)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-streamed end-to-end leg (ncu runs)")
    ap.add_argument("--small", action="store_true", help="debug: T=2048 x N=16384 panel")
    ap.add_argument("--debug-no-comm", action="store_true",
                    help="diagnostic only (results are per-shard, NOT a fit of the panel): N ranks fit their shards without the all-reduce")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
