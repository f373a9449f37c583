"""Multi-GPU plumbing for the sharded candidate grid.

One process per GPU.  Rank r sweeps rows [r·M/R, (r+1)·M/R) of the grid with ``global_offset`` set, so every
rank's Best carries a GLOBAL index; the only exchange step is the argmax: one all-gather of (value, index, mu, std)
per rank, then every rank picks the maximum value / lowest index — i.e. exactly ``np.argmin(-values)`` (first maximal
index) over the concatenated grid.  32 bytes per rank over NCCL/NVLink; latency only.
Works with the ``nccl`` backend (CUDA tensors) and with ``gloo`` (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .gp import Best

def shard_rows(M: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row block [lo, hi) of an M-row grid owned by `rank` (first M % world ranks get one extra)."""
    base, rem = divmod(M, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class RankFailure(RuntimeError):
    """A rank's local sweep failed; raised on EVERY rank after the exchange so nobody is left waiting in a collective."""


def global_argmax(best: Best | None, group=None, device=None, failed: str | None = None) -> Best:
    """Combine per-rank Bests into the global first-index argmax with ONE collective: an all-gather of 5 doubles per rank
    (value, index, mu, std, failed flag); every rank then takes the maximum value, lowest global index among equals —
    exactly ``np.argmin(-values)`` over the concatenated grid.  Indices up to 2^53 are exact in a double.
    ``failed``: a rank whose local sweep raised still JOINS the collective (with −inf and the flag set) and all ranks raise
    ``RankFailure`` together — a rank-local CUDA / not-PD error must not leave the others blocked in the all-gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if failed:
            raise RankFailure(failed)
        return best
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    if failed or best is None:
        mine = torch.tensor([float("-inf"), float(2 ** 52), 0.0, 0.0, 1.0], dtype=torch.float64, device=device)
    else:
        v = best.value if best.value == best.value else float("-inf")   # NaN never wins
        mine = torch.tensor([v, float(best.index), best.mu, best.std, 0.0], dtype=torch.float64, device=device)
    allv = torch.empty(world * 5, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    rows = allv.view(world, 5).cpu().tolist()
    bad = [r for r, row in enumerate(rows) if row[4] != 0.0]
    if bad:
        raise RankFailure(f"local sweep failed on rank(s) {bad}" + (f": {failed}" if failed else ""))
    win = max(rows, key=lambda r: (r[0], -r[1]))
    return Best(float(win[0]), int(win[1]), float(win[2]), float(win[3]))
