"""CPU, world_size 2 over gloo: the N>1 path's only exchange step (kubeflow_b200.dist.global_argmax) reproduces the
single-grid first-index argmax, including an engineered tie that straddles the two ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gp_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tie, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kubeflow_b200.dist import global_argmax, shard_rows
    from kubeflow_b200.gp import Best
    X, y, Xc = O.synthetic(64, 601, 3)
    th = O.theta_of_record(3)
    kw = dict(kind="matern52", acq="ei", **th)
    full = O.suggest(X, y, Xc, **kw)
    if tie:  # copy the winner into the other rank's shard: identical rows -> identical values -> lowest index must win
        lo0, hi0 = shard_rows(601, 0, world)
        j = (hi0 + 5) if full["index"] < hi0 else 3
        Xc = Xc.copy()
        Xc[j] = Xc[full["index"]]
        full = O.suggest(X, y, Xc, **kw)
    lo, hi = shard_rows(601, rank, world)
    part = O.suggest(X, y, Xc[lo:hi], **kw)
    mine = Best(part["value"], lo + part["index"], float(part["mu"][part["index"]]), float(part["std"][part["index"]]))
    win = global_argmax(mine)
    q.put((rank, win.index, win.value, full["index"], full["value"], win.mu, float(full["mu"][full["index"]])))
    dist.barrier()
    dist.destroy_process_group()


def _run(tie):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, tie, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, wi, wv, fi, fv, wmu, fmu in out:
        assert wi == fi, (rank, wi, fi)
        assert abs(wv - fv) < 1e-12 and abs(wmu - fmu) < 1e-9
    assert out[0][1:] == out[1][1:]      # every rank returns the same winner


def test_global_argmax_two_ranks():
    _run(tie=False)


def test_global_argmax_tie_across_ranks_lowest_index_wins():
    _run(tie=True)


def test_shard_rows_partition():
    from kubeflow_b200.dist import shard_rows
    for M, R in ((16_777_216, 8), (601, 2), (7, 4), (3, 8)):
        blocks = [shard_rows(M, r, R) for r in range(R)]
        assert blocks[0][0] == 0 and blocks[-1][1] == M
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(R - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _worker_fail(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kubeflow_b200.dist import RankFailure, global_argmax
    from kubeflow_b200.gp import Best
    try:
        if rank == 1:
            global_argmax(None, failed="KboError: boom")      # this rank's sweep raised: it still joins the exchange
        else:
            global_argmax(Best(1.0, 3, 0.0, 1.0))
        q.put((rank, "no error"))
    except RankFailure as e:
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_a_failed_rank_joins_the_exchange_and_every_rank_raises():
    """Advisor finding: a rank-local failure before the all-gather left the other ranks blocked forever.  The failing rank
    now contributes (−inf, flag) and all ranks raise RankFailure together."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_fail, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert "rank(s) [1]" in out[0] and "rank(s) [1]" in out[1] and "boom" in out[1]
