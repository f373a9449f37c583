"""Multi-GPU plumbing for the sharded candidate grid.

One process per GPU.  Rank r sweeps rows [r·M/R, (r+1)·M/R) of the grid with ``global_offset`` set, so every
rank's Best carries a GLOBAL index; the only exchange step is the argmax: an all-reduce MAX of the fp64 value,
then an all-reduce MIN of the index among the ranks holding that maximum — i.e. exactly
``np.argmin(-values)`` (first maximal index) over the concatenated grid.  16 bytes over NCCL/NVLink; latency only.
Works with the ``nccl`` backend (CUDA tensors) and with ``gloo`` (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .gp import Best

_I64_MAX = (1 << 63) - 1


def shard_rows(M: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row block [lo, hi) of an M-row grid owned by `rank` (first M % world ranks get one extra)."""
    base, rem = divmod(M, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_argmax(best: Best, group=None, device=None) -> Best:
    """Combine per-rank Bests into the global first-index argmax.  Returns the winner on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return best
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    v = best.value if best.value == best.value else float("-inf")   # NaN never wins
    val = torch.tensor([v], dtype=torch.float64, device=device)
    vmax = val.clone()
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX, group=group)
    idx = torch.tensor([best.index if v == float(vmax.item()) else _I64_MAX], dtype=torch.int64, device=device)
    dist.all_reduce(idx, op=dist.ReduceOp.MIN, group=group)
    win = int(idx.item())
    # the winner's posterior travels with a third tiny reduce (only the owner contributes non-zero)
    own = 1.0 if (best.index == win and v == float(vmax.item())) else 0.0
    post = torch.tensor([best.mu * own, best.std * own, own], dtype=torch.float64, device=device)
    dist.all_reduce(post, op=dist.ReduceOp.SUM, group=group)
    n = max(float(post[2].item()), 1.0)
    return Best(float(vmax.item()), win, float(post[0].item()) / n, float(post[1].item()) / n)
