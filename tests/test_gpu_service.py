"""GPU: algorithm `bayesianoptimization` end to end through a real in-process grpc.server — the reference-facing call
(katib-controller -> /api.v1.beta1.Suggestion/GetSuggestions) down to libkbo and back."""
import grpc
import numpy as np
import pytest

from kubeflow_b200.suggestion import api_pb as api
from kubeflow_b200.suggestion.server import SuggestionStub, serve
from kubeflow_b200.suggestion.service import DispatchService, RandomService, SkoptService
from tests.test_grpc_service import add_trial, make_experiment

pytestmark = pytest.mark.gpu


def _f(vals):   # a smooth objective with its minimum inside the box
    return (np.log10(vals["x0"]) + 1.5) ** 2 + (vals["x1"] - 0.3) ** 2 + ((vals["x2"] - 14.0) / 5) ** 2 + ((vals["x3"] - 2.0) / 2.5) ** 2


@pytest.mark.parametrize("objective", [api.MINIMIZE, api.MAXIMIZE])
def test_bayesianoptimization_over_grpc(objective):
    sk = SkoptService()
    server, port = serve(DispatchService([sk, RandomService()]), port=0, host="127.0.0.1")
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = SuggestionStub(ch)
    settings = {"base_estimator": "GP", "n_initial_points": 8, "acq_func": "EI", "acq_optimizer": "sampling", "random_state": 3,
                "n_points": 20000}
    exp = make_experiment("bayesianoptimization", settings, objective=objective, name=f"bo-{objective}")
    stub.ValidateAlgorithmSettings(api.ValidateAlgorithmSettingsRequest(experiment=exp))
    req = api.GetSuggestionsRequest(experiment=exp, current_request_number=2)
    sign = 1.0 if objective == api.MINIMIZE else -1.0
    losses = []
    for _ in range(12):                                 # 24 trials: 8 random, then GP-EI with constant-liar pairs
        reply = stub.GetSuggestions(req)
        assert len(reply.parameter_assignments) == 2
        for pa in reply.parameter_assignments:
            vals = {a.name: float(a.value) for a in pa.assignments}
            assert 0.01 <= vals["x0"] <= 0.1 and -1 <= vals["x1"] <= 1 and 10 <= vals["x2"] <= 20 and 0 <= vals["x3"] <= 5
            loss = _f(vals)
            losses.append(loss)
            add_trial(req, f"t{len(losses)}", vals, sign * loss)
    # the model-based phase must beat the random phase on this smooth bowl
    assert min(losses[8:]) < min(losses[:8])
    assert np.mean(sorted(losses[8:])[:4]) < np.mean(sorted(losses[:8])[:4])
    # steady state: each request drops the previous lie, appends the two finished trials, then appends the new lie — no refit
    assert sk._services[f"bo-{objective}"].skopt_optimizer.last_fit == "append"
    assert sk.last_ingest == "scan"                     # over gRPC the trials come through the wire scan, not the message walk
    ch.close()
    server.stop(0)


def test_optimizer_matches_oracle_argmax_on_transformed_space():
    """kubeflow_b200.Optimizer.ask == oracle argmax over the same sampled candidates (same RNG stream)."""
    from kubeflow_b200.optimizer import Optimizer
    from kubeflow_b200.space import Integer, Real, Space
    from oracle import gp_oracle as O
    sp = Space([Real(0.0, 2.0), Real(-1.0, 1.0), Integer(1, 9)])
    opt = Optimizer(sp, n_initial_points=5, acq_func="EI", random_state=11, n_points=5000, kernel="matern52", noise=1e-3,
                    candidate_backend="numpy")
    r = np.random.default_rng(0)
    pts = [[float(r.uniform(0, 2)), float(r.uniform(-1, 1)), int(r.integers(1, 10))] for _ in range(20)]
    ys = [(p[0] - 1.2) ** 2 + p[1] ** 2 + 0.05 * (p[2] - 4) ** 2 for p in pts]
    opt.tell(pts, ys)
    rng_copy = np.random.default_rng(11)
    x = opt.ask()
    cand = sp.rvs_transformed(5000, rng_copy, np.float32).astype(np.float64)
    ref = O.suggest(sp.transform(pts), np.asarray(ys), cand, kind="matern52", acq="ei", length_scale=0.3 * np.sqrt(3), amplitude=1.0,
                    noise=1e-3)
    assert opt.last_best.index == ref["index"] and abs(opt.last_best.value - ref["value"]) < 1e-7
    assert x == sp.inverse_transform(cand[ref["index"]:ref["index"] + 1])[0]


def test_cmaes_over_grpc():
    """Katib algorithm `cmaes` on the GPU sampler: three full generations through the gRPC surface improve a bowl."""
    from kubeflow_b200.suggestion.cmaes_service import CmaesService
    server, port = serve(DispatchService([CmaesService()]), port=0, host="127.0.0.1")
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = SuggestionStub(ch)
    exp = make_experiment("cmaes", {"random_state": 2, "popsize": 12, "sigma": 0.25}, name="cma-1")
    stub.ValidateAlgorithmSettings(api.ValidateAlgorithmSettingsRequest(experiment=exp))
    req = api.GetSuggestionsRequest(experiment=exp, current_request_number=6)
    gen_means = []
    losses = []
    for call in range(12):                     # 6 generations of 12, two calls each
        reply = stub.GetSuggestions(req)
        assert len(reply.parameter_assignments) == 6
        for pa in reply.parameter_assignments:
            vals = {a.name: float(a.value) for a in pa.assignments}
            assert 0.01 <= vals["x0"] <= 0.1 and -1 <= vals["x1"] <= 1 and 10 <= vals["x2"] <= 20 and 0 <= vals["x3"] <= 5
            losses.append(_f(vals))
            add_trial(req, f"t{len(losses)}", vals, losses[-1])
        if call % 2:
            gen_means.append(np.mean(losses[-12:]))
    assert gen_means[-1] < gen_means[0]
    with pytest.raises(grpc.RpcError) as ei:
        bad = make_experiment("cmaes", {"popsize": 2})
        stub.ValidateAlgorithmSettings(api.ValidateAlgorithmSettingsRequest(experiment=bad))
    assert ei.value.code() == grpc.StatusCode.INVALID_ARGUMENT
    ch.close()
    server.stop(0)


def test_optimizer_device_candidates_match_oracle():
    """Default candidate backend: sampled on the device with a seeded torch.Generator; the same stream re-drawn here must
    give the oracle the same winner."""
    import torch
    from kubeflow_b200.optimizer import Optimizer
    from kubeflow_b200.space import Categorical, Integer, Real, Space
    from oracle import gp_oracle as O
    sp = Space([Real(0.0, 2.0), Integer(1, 9), Categorical(["a", "b", "c"])])
    opt = Optimizer(sp, n_initial_points=4, acq_func="LCB", random_state=5, n_points=4096, kernel="rbf", noise=1e-3)
    r = np.random.default_rng(1)
    pts = [[float(r.uniform(0, 2)), int(r.integers(1, 10)), ["a", "b", "c"][int(r.integers(0, 3))]] for _ in range(16)]
    ys = [(p[0] - 1.0) ** 2 + 0.1 * p[1] + (0.5 if p[2] == "b" else 0.0) for p in pts]
    opt.tell(pts, ys)
    x = opt.ask()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    cand = sp.rvs_transformed_torch(4096, g, torch.device("cuda", 0)).cpu().numpy().astype(np.float64)
    assert np.allclose(cand[:, 2:5].sum(1), 1.0) and np.allclose(cand[:, 1] * 8, np.round(cand[:, 1] * 8), atol=1e-5)
    ref = O.suggest(sp.transform(pts), np.asarray(ys), cand, kind="rbf", acq="lcb", length_scale=0.3 * np.sqrt(5), amplitude=1.0, noise=1e-3)
    assert opt.last_best.index == ref["index"] and abs(opt.last_best.value - ref["value"]) < 1e-7
    assert x == sp.inverse_transform(cand[ref["index"]:ref["index"] + 1])[0]


def test_theta_grid_picks_higher_lml_length_scale():
    """SURVEY.md §8(f)1 (first step): length scale chosen among a grid by the GPU-computed log-marginal likelihood."""
    from kubeflow_b200.gp import GPEngine
    from kubeflow_b200.optimizer import Optimizer
    from kubeflow_b200.space import Real, Space
    from oracle import gp_oracle as O
    r = np.random.default_rng(0)
    sp = Space([Real(0.0, 1.0), Real(0.0, 1.0)])
    pts = r.random((60, 2)).tolist()
    ys = [float(np.sin(25 * p[0]) + np.cos(19 * p[1])) for p in pts]        # wiggly: the default ℓ = 0.3·√2 is far too long
    opt = Optimizer(sp, n_initial_points=5, acq_func="EI", random_state=1, n_points=2048, theta_grid=7, noise=1e-4)
    opt.tell(pts, ys)
    opt.ask()
    chosen = float(opt._engine.length_scale[0])
    base = 0.3 * np.sqrt(2)
    assert chosen < base * 0.6
    lml = lambda ls: O.gp_fit(np.asarray(pts), np.asarray(ys), kind="matern52", length_scale=ls, noise=1e-4)["lml"]
    assert lml(chosen) > lml(base) + 1.0
    grid = base * np.geomspace(0.25, 4.0, 7)
    assert abs(chosen - grid[int(np.argmax([lml(g) for g in grid]))]) < 1e-12      # same pick as the oracle's LML over the same grid


def test_theta_search_improves_lml_over_default():
    """Random (length scale, noise) search by GPU LML: never worse than the default θ, and clearly better on a noisy wiggly target."""
    from kubeflow_b200.optimizer import Optimizer
    from kubeflow_b200.space import Real, Space
    from oracle import gp_oracle as O
    r = np.random.default_rng(3)
    sp = Space([Real(0.0, 1.0), Real(0.0, 1.0), Real(0.0, 1.0)])
    pts = r.random((120, 3)).tolist()
    ys = [float(np.sin(9 * p[0]) * np.cos(7 * p[1]) + 0.5 * p[2] + 0.2 * r.standard_normal()) for p in pts]
    opt = Optimizer(sp, n_initial_points=5, acq_func="EI", random_state=2, n_points=1024, theta_search=24)
    opt.tell(pts, ys)
    opt.ask()
    th = opt.last_theta
    base = O.gp_fit(np.asarray(pts), np.asarray(ys), kind="matern52", length_scale=0.3 * np.sqrt(3), noise=1e-3)["lml"]
    chosen = O.gp_fit(np.asarray(pts), np.asarray(ys), kind="matern52", length_scale=th["length_scale"], noise=th["noise"])["lml"]
    assert abs(chosen - th["lml"]) < 1e-6 * abs(chosen)          # the GPU LML that drove the choice equals the oracle's at that θ
    assert chosen > base + 5.0 and th["noise"] > 1e-3            # the data are noisy: a larger noise level wins


def test_theta_lbfgs_polish_uses_gpu_gradients():
    """L-BFGS-B over (log ℓ_1..D, log noise) on device LML + gradient: ends at a stationary point whose LML the oracle confirms,
    above both the default θ and the random-search pick it starts from."""
    from kubeflow_b200.optimizer import Optimizer
    from kubeflow_b200.space import Real, Space
    from oracle import gp_oracle as O
    r = np.random.default_rng(5)
    sp = Space([Real(0.0, 1.0), Real(0.0, 1.0), Real(0.0, 1.0)])
    pts = r.random((150, 3)).tolist()
    ys = [float(np.sin(11 * p[0]) + 0.3 * np.cos(2 * p[1]) + 0.1 * r.standard_normal()) for p in pts]   # x0 matters most, x2 not at all
    Xn, yn = np.asarray(pts), np.asarray(ys)
    base = O.gp_fit(Xn, yn, kind="matern52", length_scale=0.3 * np.sqrt(3), noise=1e-3)["lml"]
    o1 = Optimizer(sp, n_initial_points=5, random_state=2, n_points=512, theta_search=12)
    o1.tell(pts, ys); o1.ask()
    o2 = Optimizer(sp, n_initial_points=5, random_state=2, n_points=512, theta_search=12, theta_fit="lbfgs", ard=True, theta_fit_maxiter=40)
    o2.tell(pts, ys); o2.ask()
    t = o2.last_theta
    ref_lml, ref_g = O.lml_and_grad(Xn, yn, kind="matern52", length_scale=t["length_scale"], noise=t["noise"])
    assert abs(ref_lml - t["lml"]) < 1e-6 * abs(ref_lml)
    assert t["lml"] > o1.last_theta["lml"] + 1.0 > base + 1.0
    assert t["length_scale"][0] < t["length_scale"][2]                   # ARD found the relevant dimension
    assert np.abs(ref_g[1:]).max() < 2.0                                 # near-stationary in (noise, ℓ) after ≤ 40 iterations


def test_concurrent_requests_and_lru_eviction():
    """Four client threads hit the server at once for different experiments (handlers run on a thread pool; GPU work is
    serialised by the service lock); with max_experiments = 2 older experiments' engines are evicted and rebuilt on demand."""
    import threading
    svc = SkoptService(max_experiments=2)
    server, port = serve(DispatchService([svc]), port=0, host="127.0.0.1")
    errors, results = [], {}

    def client(k):
        try:
            ch = grpc.insecure_channel(f"127.0.0.1:{port}")
            stub = SuggestionStub(ch)
            exp = make_experiment("bayesianoptimization", {"n_initial_points": 2, "acq_func": "EI", "random_state": k, "n_points": 2000}, name=f"c{k}")
            req = api.GetSuggestionsRequest(experiment=exp, current_request_number=1)
            for i in range(6):
                rep = stub.GetSuggestions(req)
                vals = {a.name: float(a.value) for a in rep.parameter_assignments[0].assignments}
                add_trial(req, f"c{k}-t{i}", vals, _f(vals))
            results[k] = len(req.trials)
            ch.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=client, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not errors, errors
    assert results == {0: 6, 1: 6, 2: 6, 3: 6}
    assert len(svc._services) <= 2
    server.stop(0)
