"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the SPMD suggestion service with the real engine — the sharded
sweep's global winner is the better of the two shard winners, both ranks stay in lockstep, replies are valid points."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, pg_port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(pg_port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import grpc
    from kubeflow_b200.suggestion import api_pb as api
    from kubeflow_b200.suggestion.server import SuggestionStub, serve
    from kubeflow_b200.suggestion.service import DispatchService, RandomService, SkoptService
    from kubeflow_b200.suggestion.spmd import SpmdServicer
    from tests.test_grpc_service import add_trial, make_experiment
    sk = SkoptService({"device": rank, "shard": True})
    spmd = SpmdServicer(DispatchService([sk, RandomService()]))
    result = {"rank": rank}
    if rank == 0:
        server, port = serve(spmd, port=0, host="127.0.0.1")
        ch = grpc.insecure_channel(f"127.0.0.1:{port}")
        stub = SuggestionStub(ch)
        exp = make_experiment("bayesianoptimization", {"n_initial_points": 4, "acq_func": "EI", "random_state": 5, "n_points": 200001}, name="spmd")
        req = api.GetSuggestionsRequest(experiment=exp, current_request_number=2)
        rng = np.random.default_rng(0)
        n, replies = 0, []
        for _ in range(6):
            rep = stub.GetSuggestions(req)
            pts = [{a.name: float(a.value) for a in pa.assignments} for pa in rep.parameter_assignments]
            replies.append(pts)
            for v in pts:
                n += 1
                add_trial(req, f"t{n}", v, float((v["x1"] - 0.3) ** 2 + ((v["x2"] - 14.0) / 5) ** 2 + 0.01 * rng.standard_normal()))
        result["replies"] = replies
        ch.close()
        server.stop(0)
        spmd.stop()
    else:
        spmd.worker_loop()
    opt = sk._services["spmd"].skopt_optimizer
    result.update(calls=spmd.calls, Xi=np.asarray(opt.Xi).tolist(), yi=list(opt.yi), last_best=(opt.last_best.index, opt.last_best.value),
                  local_best=(opt.last_local_best.index, opt.last_local_best.value), last_fit=opt.last_fit)
    q.put(result)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_spmd_service_two_gpus():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=600) for _ in ps], key=lambda d: d["rank"])
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = out
    assert r0["calls"] == r1["calls"] == 6
    assert r0["Xi"] == r1["Xi"] and r0["yi"] == r1["yi"] and len(r0["yi"]) == 10
    assert r0["last_best"] == r1["last_best"]
    assert r0["last_best"] == max([r0["local_best"], r1["local_best"]], key=lambda b: (b[1], -b[0]))
    assert 0 <= r0["local_best"][0] < 100001 <= r1["local_best"][0] < 200001
    assert r0["last_fit"] == r1["last_fit"] == "append"
    for pts in r0["replies"]:
        assert len(pts) == 2 and all(-1 <= v["x1"] <= 1 and 10 <= v["x2"] <= 20 for v in pts)


def _comm_worker(rank, world, uid, q):
    """C-ABI exchange: each rank's engine sweeps its shard and kbo_allreduce_argmax (ncclAllGather + reduce) gives the global winner."""
    from kubeflow_b200.dist import shard_rows
    from kubeflow_b200.gp import GPEngine
    from oracle import gp_oracle as O
    torch.cuda.set_device(rank)
    N, M, D = 1500, 40001, 7
    X, y, Xc = O.synthetic(N, M, D)
    th = O.theta_of_record(D)
    if rank == 0:
        uid_bytes = GPEngine.comm_unique_id()
        uid.put(uid_bytes)
    else:
        uid_bytes = uid.get(timeout=120)
    eng = GPEngine(rank, kernel="matern52", acq="ei", var_mode="tc", **th)
    eng.comm_init(world, rank, uid_bytes)
    assert eng.comm_size() == world
    eng.tell(X, y)
    lo, hi = shard_rows(M, rank, world)
    local = eng.ask(Xc[lo:hi], global_offset=lo)
    glob = eng.ask(Xc[lo:hi], global_offset=lo, allreduce=True)
    full = eng.ask(Xc) if rank == 0 else None
    q.put((rank, (local.index, local.value), (glob.index, glob.value, glob.mu, glob.std), (full.index, full.value) if full else None))
    eng.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_c_abi_allreduce_argmax_two_gpus():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, uid = ctx.Queue(), ctx.Queue()
    ps = [ctx.Process(target=_comm_worker, args=(r, 2, uid, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=600) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, l0, g0, full), (_, l1, g1, _) = out
    assert g0 == g1                                                  # every rank holds the same global result
    assert g0[:2] == max([l0, l1], key=lambda b: (b[1], -b[0]))       # the better shard winner, lowest index on ties
    assert g0[:2] == full                                            # = the single-GPU sweep of the whole grid
