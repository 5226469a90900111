#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the column-sharded closed-form and
three-pass fits must match the single-process CPU oracle on the full panel to 1e-9, and every rank
must hold bit-identical forecasts (fixed rank-order reduction)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
import uni_b200 as u
from oracle import tprf_oracle as to
from uni_b200.dist import destroy_comm, init_comm, shard_range

u.init(local)


def bcast(b):
    t = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


init_comm(rank, world, bcast)
import ctypes as _C
from uni_b200._lib import lib as _lib
_p2p = _C.c_int32(0); _lib.um_comm_transport(_C.byref(_p2p))
print(f"rank {rank}: exchange transport = {'peer memory (CUDA IPC)' if _p2p.value else 'NCCL'}", flush=True)
ok = True
for (T, N, L) in ((512, 4096, 4), (300, 1000, 3), (1024, 8192, 8), (640, 6000, 13), (2048, 16384, 20)):   # L > 8: one sweep per group of 8 + operator finish
    rng = np.random.default_rng(7)
    X = rng.standard_normal((T, N)) * 2 + 0.3; y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    c0, c1 = shard_range(N, rank, world)
    dy, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(Z)
    dX = u.Mat.from_numpy(np.ascontiguousarray(X[:, c0:c1]))
    c = u.tprfClosedForm(dy, dX, dZ, n_total=N)
    it = u.t3prf(dy, dX, dZ, n_total=N)
    oc = to.tprf_closed_form_gram(y, X, Z)
    oi = to.t3prf(y, X, Z)
    tol = lambda a, b: np.max(np.abs(a - b) / (1e-12 + 1e-9 * np.maximum(np.abs(a), np.abs(b))))
    e1 = tol(c.forecasts.to_numpy(), oc.forecasts); e2 = tol(it.forecasts.to_numpy(), oi.forecasts)
    e3 = tol(c.alpha.to_numpy(), oc.alpha[c0:c1]) if c1 > c0 else 0.0
    e4 = tol(it.phi.to_numpy(), oi.phi[c0:c1]) if c1 > c0 else 0.0
    f = torch.from_numpy(c.forecasts.to_numpy()).cuda()
    g = [torch.empty_like(f) for _ in range(world)]
    dist.all_gather(g, f)
    same = all(torch.equal(g[0], x) for x in g)
    good = e1 <= 1 and e2 <= 1 and e3 <= 10 and e4 <= 10 and same
    ok = ok and good
    print(f"rank {rank} T={T} N={N} L={L} shard=[{c0},{c1}) closed {e1:.3g} iter {e2:.3g} alpha {e3:.3g} phi {e4:.3g} "
          f"(x tolerance)  ranks bit-identical: {same}  R2 {c.rSquared:.6g} vs {oc.r_squared:.6g}", flush=True)
# many back-to-back exchanges: slot parity reuse, counters, and the per-exchange cost
import time
buf = u.MatD.randn(1, 73809, seed=rank)
ref_sum = None
for trial in range(3):
    parts = [u.MatD.randn(1, 73809, seed=r).to_numpy() for r in range(world)]
    want = parts[0].copy()
    for r in range(1, world):
        want = want + parts[r]
    b2 = buf.copy()
    _lib.um_comm_allreduce_sum_f64(_C.c_void_p(b2._buf.ptr), 73809)
    same = np.array_equal(b2.to_numpy(), want)
    ok = ok and same
    if not same:
        print(f"rank {rank}: all-reduce trial {trial} WRONG (max diff {np.max(np.abs(b2.to_numpy() - want))})", flush=True)
u.sync(); dist.barrier(); t0 = time.perf_counter()
b3 = buf.copy()
for _ in range(200):
    _lib.um_comm_allreduce_sum_f64(_C.c_void_p(b3._buf.ptr), 73809)
u.sync()
print(f"rank {rank}: 200 exchanges of 73809 doubles: {(time.perf_counter() - t0) / 200 * 1e3:.4f} ms each", flush=True)
dist.barrier()
destroy_comm()
dist.destroy_process_group()
print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)
