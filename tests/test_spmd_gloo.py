"""CPU, world_size 2 over gloo: the SPMD suggestion service (suggestion/spmd.py) — rank 0 serves gRPC, every request is
broadcast and run on both ranks with the candidate grid sharded.  The GPU engine is replaced by a test-only stand-in built
on the oracle (tests may use oracle/; the product path has no CPU engine), so what is tested is the host logic: lockstep
optimizers, per-rank candidate shards, the argmax exchange, the winner's row broadcast, constant-liar batches across ranks."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleEngine:
    """Same surface as kubeflow_b200.gp.GPEngine (tell / append / rebase / room / ask), numpy + oracle underneath."""

    def __init__(self):
        self.kernel, self.acq, self.var_mode, self.normalize_y = "matern52", "ei", "auto", True
        self.length_scale, self.amplitude, self.noise, self.xi, self.kappa = np.ones(1), 1.0, 1e-3, 0.01, 1.96
        self.X = self.y = None
        self.log = []

    def tell(self, X, y):
        self.X, self.y = np.array(X, dtype=np.float64), np.array(y, dtype=np.float64)
        self.log.append("fit")

    def room(self):
        return (-len(self.y)) % 64

    def append(self, x, y):
        self.X, self.y = np.vstack([self.X, np.asarray(x)[None, :]]), np.append(self.y, y)
        self.log.append("append")

    def rebase(self, n, y=None):
        self.X = self.X[:n]
        self.y = np.array(y[:n], dtype=np.float64) if y is not None else self.y[:n]
        self.log.append("rebase")

    def ask(self, cand, global_offset=0):
        from kubeflow_b200.gp import Best
        from oracle import gp_oracle as O
        r = O.suggest(self.X, self.y, np.asarray(cand, dtype=np.float64), kind=self.kernel, acq=self.acq, length_scale=self.length_scale,
                      amplitude=self.amplitude, noise=self.noise, xi=self.xi, kappa=self.kappa)
        i = r["index"]
        return Best(r["value"], global_offset + i, float(r["mu"][i]), float(r["std"][i]))

    def close(self):
        pass


def _worker(rank, world, pg_port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(pg_port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import grpc
    from kubeflow_b200.suggestion import api_pb as api
    from kubeflow_b200.suggestion.server import SuggestionStub, serve
    from kubeflow_b200.suggestion.service import DispatchService, RandomService, SkoptService
    from kubeflow_b200.suggestion.spmd import SpmdServicer
    from tests.test_grpc_service import add_trial, make_experiment
    eng = OracleEngine()
    sk = SkoptService({"engine": eng, "candidate_backend": "numpy", "shard": True})
    spmd = SpmdServicer(DispatchService([sk, RandomService()]))
    result = {"rank": rank}
    if rank == 0:
        server, port = serve(spmd, port=0, host="127.0.0.1")
        ch = grpc.insecure_channel(f"127.0.0.1:{port}")
        stub = SuggestionStub(ch)
        exp = make_experiment("bayesianoptimization", {"n_initial_points": 4, "acq_func": "EI", "random_state": 5, "n_points": 601}, name="spmd")
        req = api.GetSuggestionsRequest(experiment=exp, current_request_number=2)
        rng = np.random.default_rng(0)
        n = 0
        replies = []
        for _ in range(5):                      # 2 random rounds (n_initial_points=4), then 3 model-based rounds of 2 points
            rep = stub.GetSuggestions(req)
            pts = [{a.name: float(a.value) for a in pa.assignments} for pa in rep.parameter_assignments]
            replies.append(pts)
            for v in pts:
                n += 1
                add_trial(req, f"t{n}", v, float(sum(v.values()) + 0.01 * rng.standard_normal()))
        # an algorithm with nothing to shard stays on rank 0 (no broadcast, the worker never sees it)
        rexp = make_experiment("random", {"random_state": 1}, name="spmd-random")
        assert len(stub.GetSuggestions(api.GetSuggestionsRequest(experiment=rexp, current_request_number=3)).parameter_assignments) == 3
        result["replies"] = replies
        ch.close()
        server.stop(0)
        spmd.stop()
    else:
        spmd.worker_loop()
    opt = sk._services["spmd"].skopt_optimizer
    result.update(calls=spmd.calls, Xi=np.asarray(opt.Xi).tolist(), yi=list(opt.yi), last_best=(opt.last_best.index, opt.last_best.value),
                  local_best=(opt.last_local_best.index, opt.last_local_best.value), log=list(eng.log))
    q.put(result)
    dist.barrier()
    dist.destroy_process_group()


def test_spmd_service_two_ranks_lockstep_and_sharded_argmax():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=300) for _ in ps], key=lambda d: d["rank"])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = out
    assert r0["calls"] == r1["calls"] == 5                       # the `random` request was not broadcast
    assert r0["Xi"] == r1["Xi"] and r0["yi"] == r1["yi"] and len(r0["yi"]) == 8   # lockstep: identical tells on both ranks
    assert r0["last_best"] == r1["last_best"]                    # the same global winner everywhere …
    lb = [r0["local_best"], r1["local_best"]]
    assert r0["last_best"] == max(lb, key=lambda b: (b[1], -b[0]))   # … which is the better of the two shard winners
    assert 0 <= r0["local_best"][0] < 301 <= r1["local_best"][0] < 601   # rank r swept rows [lo_r, hi_r) of the 601-candidate grid
    assert r0["log"] == r1["log"] and "append" in r0["log"]      # constant lies appended on both ranks alike
    for pts in r0["replies"]:
        assert len(pts) == 2
        for v in pts:
            assert 0.01 <= v["x0"] <= 0.1 and -1 <= v["x1"] <= 1 and 10 <= v["x2"] <= 20 and 0 <= v["x3"] <= 5
