"""Per-call latency of GetSuggestions at the sizes Katib experiments usually have (tens to hundreds of trials, skopt's default
n_points = 10000), through a real in-process grpc.server; optional cProfile of the servicer for the host-side hot spots."""
import cProfile
import json
import pstats
import sys
import time

import grpc
import numpy as np

sys.path.insert(0, ".")
from kubeflow_b200.suggestion import api_pb as api  # noqa: E402
from kubeflow_b200.suggestion.server import SuggestionStub, serve  # noqa: E402
from kubeflow_b200.suggestion.service import SkoptService  # noqa: E402


def experiment(name, D, n_points):
    ex = api.Experiment()
    ex.name = name
    ex.spec.objective.type = api.MINIMIZE
    ex.spec.objective.objective_metric_name = "loss"
    ex.spec.algorithm.algorithm_name = "bayesianoptimization"
    for k, v in {"n_initial_points": 5, "acq_func": "EI", "random_state": 1, "n_points": n_points}.items():
        s = ex.spec.algorithm.algorithm_settings.add()
        s.name, s.value = k, str(v)
    for d in range(D):
        p = ex.spec.parameter_specs.parameters.add()
        p.name, p.parameter_type = f"x{d}", api.DOUBLE
        p.feasible_space.min, p.feasible_space.max = "0", "1"
    return ex


def main():
    svc = SkoptService({"device": 0})
    server, port = serve(svc, port=0, host="127.0.0.1")
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = SuggestionStub(ch)
    out = []
    for N, D, M, k in ((30, 4, 10000, 1), (30, 4, 10000, 3), (300, 16, 10000, 1), (300, 16, 10000, 3), (1000, 16, 65536, 1)):
        rng = np.random.default_rng(N + D)
        req = api.GetSuggestionsRequest(experiment=experiment(f"lat-{N}-{D}-{M}-{k}", D, M), current_request_number=k)

        def add(i):
            t = req.trials.add()
            t.name = f"t{i}"
            t.spec.objective.objective_metric_name = "loss"
            t.status.condition = api.SUCCEEDED
            x = rng.random(D)
            for d in range(D):
                a = t.spec.parameter_assignments.assignments.add()
                a.name, a.value = f"x{d}", repr(float(x[d]))
            m = t.status.observation.metrics.add()
            m.name, m.value = "loss", repr(float(np.sin(3 * x.sum()) + 0.1 * rng.standard_normal()))

        for i in range(N):
            add(i)
        calls = []
        for c in range(12):
            t0 = time.perf_counter()
            stub.GetSuggestions(req)
            calls.append((time.perf_counter() - t0) * 1e3)
            for j in range(k):
                add(N + c * k + j)
        out.append({"N": N, "D": D, "n_points": M, "request_number": k, "first_ms": calls[0], "steady_median_ms": float(np.median(calls[2:])),
                    "steady_min_ms": float(np.min(calls[2:])), "last_fit": svc._services[req.experiment.name].skopt_optimizer.last_fit})
    print(json.dumps(out))
    if "--profile" in sys.argv:
        data = req.SerializeToString()
        from kubeflow_b200.suggestion.ingest import LazyRequest
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(20):
            svc.GetSuggestions(LazyRequest.FromString(data), None)
        pr.disable()
        pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
    ch.close()
    server.stop(0)


if __name__ == "__main__":
    main()
