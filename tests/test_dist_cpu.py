"""world_size-2 gloo test of the sharded 3PRF exchange: every rank builds the shard-local
accumulators for its column block (shard_range), ONE packed sum-all-reduce, identical finish on
every rank -- must reproduce the single-process fit.  CPU only; the device
kernels produce the same packed layout (uni_b200/csrc/tprf.cu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tprf_oracle as to
from uni_b200.dist import packed_len, shard_range

T, N, L = 96, 333, 3


def _panel():
    rng = np.random.default_rng(21)
    return rng.standard_normal((T, 1)), rng.standard_normal((T, N)) * 2 + 0.5, rng.standard_normal((T, L))


def _pack(p):
    px = np.zeros((T, 8)); px[:, :L] = p["PX"]
    sw = np.zeros(8); sw[:L] = p["sw"]
    sww = np.zeros((8, 8)); sww[:L, :L] = p["SWW"]
    smw = np.zeros(8); smw[:L] = p["smw"]
    return np.concatenate([px.ravel(), p["h0"], sw, sww.ravel(), smw, [p["smu"]]])


def _unpack(v):
    px = v[: T * 8].reshape(T, 8)[:, :L]
    h0 = v[T * 8 : T * 9]
    s = v[T * 9 :]
    return {"PX": px, "h0": h0, "sw": s[:8][:L], "SWW": s[8:72].reshape(8, 8)[:L, :L], "smw": s[72:80][:L], "smu": s[80]}


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y, X, Z = _panel()
    c0, c1 = shard_range(N, rank, world)
    parts = to.closed_form_partials(X[:, c0:c1], Z)
    packed = torch.from_numpy(_pack(parts))
    assert packed.numel() == packed_len(T)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    yhat, s, wbar = to.closed_form_from_partials(y, _unpack(packed.numpy()), N)
    alpha = (parts["W0"] - wbar) @ s
    np.savez(os.path.join(out, f"r{rank}.npz"), yhat=yhat, alpha=alpha, c0=c0, c1=c1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_sharded_fit_matches_single(tmp_path, world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    y, X, Z = _panel()
    full = to.tprf_closed_form_gram(y, X, Z)
    alpha = np.zeros((N, 1))
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_allclose(g["yhat"], full.forecasts, rtol=1e-10, atol=1e-13)
        alpha[int(g["c0"]) : int(g["c1"])] = g["alpha"]
    np.testing.assert_allclose(alpha, full.alpha, rtol=1e-9, atol=1e-13)
