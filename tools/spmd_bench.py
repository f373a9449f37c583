"""GetSuggestions through the SPMD service at cfg3 scale per GPU: N = 8192 finished trials resent as strings, n_points =
1M × world candidates sharded over `world` GPUs (python tools/spmd_bench.py [world]; spawns one process per GPU, NCCL)."""
import json
import os
import socket
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, pg_port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(pg_port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import grpc
    from kubeflow_b200.suggestion import api_pb as api
    from kubeflow_b200.suggestion.server import SuggestionStub, serve
    from kubeflow_b200.suggestion.service import SkoptService
    from kubeflow_b200.suggestion.spmd import SpmdServicer
    sk = SkoptService({"device": rank, "shard": True})
    spmd = SpmdServicer(sk)
    if rank == 0:
        N, D, M = 8192, 32, (1 << 20) * world
        X = np.random.default_rng(1234).random((N, D))
        y = np.sin(3.0 * X.sum(axis=1) / np.sqrt(D)) + 0.1 * np.random.default_rng(1235).standard_normal(N)
        server, port = serve(spmd, port=0, host="127.0.0.1")
        ch = grpc.insecure_channel(f"127.0.0.1:{port}", options=[("grpc.max_send_message_length", 1 << 28), ("grpc.max_receive_message_length", 1 << 28)])
        stub = SuggestionStub(ch)
        ex = api.Experiment()
        ex.name = "spmd-cfg3"
        ex.spec.objective.type = api.MINIMIZE
        ex.spec.objective.objective_metric_name = "loss"
        ex.spec.algorithm.algorithm_name = "bayesianoptimization"
        for k_, v_ in {"n_initial_points": 0, "acq_func": "EI", "random_state": 1, "n_points": M, "var_mode": "tc"}.items():
            st_ = ex.spec.algorithm.algorithm_settings.add()
            st_.name, st_.value = k_, str(v_)
        for d_ in range(D):
            ps_ = ex.spec.parameter_specs.parameters.add()
            ps_.name, ps_.parameter_type = f"x{d_}", api.DOUBLE
            ps_.feasible_space.min, ps_.feasible_space.max = "0", "1"
        rq = api.GetSuggestionsRequest(experiment=ex, current_request_number=1)

        def add_trial(i):
            t_ = rq.trials.add()
            t_.name = f"t{i}"
            t_.spec.objective.objective_metric_name = "loss"
            t_.status.condition = api.SUCCEEDED
            for d_ in range(D):
                a_ = t_.spec.parameter_assignments.assignments.add()
                a_.name, a_.value = f"x{d_}", repr(float(X[i % N, d_]))
            m_ = t_.status.observation.metrics.add()
            m_.name, m_.value = "loss", repr(float(y[i % N]))

        for i in range(N - 5):
            add_trial(i)
        calls = []
        for c in range(6):
            t0 = time.perf_counter()
            stub.GetSuggestions(rq)
            calls.append((time.perf_counter() - t0) * 1e3)
            add_trial(N - 5 + c)
        opt = sk._services["spmd-cfg3"].skopt_optimizer
        q.put({"world": world, "trials": N, "n_points_total": M, "request_bytes": rq.ByteSize(), "cold_call_ms": calls[0],
               "steady_call_ms": float(np.median(calls[1:])), "calls_ms": calls, "engine_update_last_call": opt.last_fit,
               "ingest_last_call": sk.last_ingest})
        ch.close()
        server.stop(0)
        spmd.stop()
    else:
        spmd.worker_loop()
    dist.barrier()
    dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp
    world = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    print(json.dumps(q.get(timeout=900)))
    for p in ps:
        p.join(timeout=120)


if __name__ == "__main__":
    main()
