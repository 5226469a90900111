"""Multi-GPU plumbing for the column-sharded 3PRF fit.

One process per GPU.  X is partitioned by predictor column into contiguous blocks; y and Z are
replicated.  The only exchange is one packed sum of T*(8+1)+81 doubles per fit, done inside
libunimat: an all-gather into mapped peer memory (CUDA IPC over NVLink) + fixed rank-order add in one
kernel, NCCL all-gather as the fallback transport.  torch.distributed is used only to hand the NCCL
unique id from rank 0 to the other ranks.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Tuple

from ._lib import check, lib


def shard_range(n_cols: int, rank: int, world: int, align: int = 64) -> Tuple[int, int]:
    """Contiguous column block [c0, c1) of rank `rank`: blocks are multiples of `align` columns
    (the sweep's strip width) while columns last, the remainder goes to the trailing ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    units = -(-n_cols // align)
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return min(u0 * align, n_cols), min(u1 * align, n_cols)


def packed_len(T: int) -> int:
    """Doubles exchanged per fit: PX[T][8] | h0[T] | sw[8] | SWW[64] | smw[8] | smu."""
    return T * 9 + 81


def init_comm(rank: int, world: int, broadcast_bytes: Callable[[bytes], bytes]) -> None:
    """`broadcast_bytes(b)` must return rank 0's `b` on every rank (e.g. via torch.distributed)."""
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        check(lib.um_comm_unique_id(buf))
    raw = broadcast_bytes(bytes(buf))
    ident = (C.c_uint8 * 128).from_buffer_copy(raw)
    check(lib.um_comm_init(rank, world, ident))


def destroy_comm() -> None:
    check(lib.um_comm_destroy())


def transport() -> str:
    """Which transport the exchange step uses on this rank: "p2p" (mapped peer memory) or "nccl"."""
    p = C.c_int32(0)
    check(lib.um_comm_transport(C.byref(p)))
    return "p2p" if p.value else "nccl"
