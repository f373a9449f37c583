"""CPU: Hyperband bracket bookkeeping through the real gRPC surface, state round-tripping through
reply.algorithm.algorithm_settings exactly as a stateless upstream service does."""
import grpc
import pytest

from kubeflow_b200.suggestion import api_pb as api
from kubeflow_b200.suggestion.hyperband import HyperbandService, bracket_plan
from kubeflow_b200.suggestion.server import SuggestionStub, serve
from tests.test_grpc_service import add_trial, make_experiment


def test_bracket_plan_matches_formula():
    # eta = 3, R = 81 (the paper's running example): s_max = 4, n = ceil(5/(s+1)·3^s), r = 81·3^-s
    plan = bracket_plan(3, 81)
    assert [s for s, _, _ in plan] == [4, 3, 2, 1, 0]
    assert [n for _, n, _ in plan] == [81, 34, 15, 8, 5]
    assert [round(r, 9) for _, _, r in plan] == [1, 3, 9, 27, 81]
    assert bracket_plan(2, 16)[0][:2] == (4, 16) and bracket_plan(4, 64)[0][:2] == (3, 64)


def _exp(settings):
    e = make_experiment("hyperband", settings, name="hb")
    p = e.spec.parameter_specs.parameters.add()
    p.name, p.parameter_type = "epochs", api.INT
    p.feasible_space.min, p.feasible_space.max = "1", "27"
    return e


def test_successive_halving_over_grpc():
    server, port = serve(HyperbandService(), port=0, host="127.0.0.1")
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = SuggestionStub(ch)
    exp = _exp({"eta": 3, "r_l": 27, "resource_name": "epochs", "random_state": 4})
    stub.ValidateAlgorithmSettings(api.ValidateAlgorithmSettingsRequest(experiment=exp))
    req = api.GetSuggestionsRequest(experiment=exp, current_request_number=27)
    sizes, resources, all_trials = [], [], 0
    for rung in range(9):
        reply = stub.GetSuggestions(req)
        pts = [{a.name: a.value for a in pa.assignments} for pa in reply.parameter_assignments]
        if not pts:
            break
        sizes.append(len(pts))
        resources.append({p["epochs"] for p in pts})
        names = []
        for p in pts:
            all_trials += 1
            name = f"t{all_trials}"
            names.append(name)
            add_trial(req, name, p, loss=float(p["x1"]) ** 2 + 1.0 / float(p["epochs"]))
        # the controller echoes the state back, with the names it gave to the rung's trials
        del req.experiment.spec.algorithm.algorithm_settings[:]
        for s in reply.algorithm.algorithm_settings:
            x = req.experiment.spec.algorithm.algorithm_settings.add()
            x.name, x.value = s.name, (",".join(names) if s.name == "bracket_trials" else s.value)
    # s_max = 3: bracket 3 = 27@1 → 9@3 → 3@9 → 1@27, then bracket 2 = ceil(4/3·9)=12@3 → 4@9 → 1@27, then bracket 1 = 6@9 ...
    assert sizes[:7] == [27, 9, 3, 1, 12, 4, 1]
    assert resources[:7] == [{"1"}, {"3"}, {"9"}, {"27"}, {"3"}, {"9"}, {"27"}]
    ch.close()
    server.stop(0)


def test_survivors_are_the_best_of_the_rung():
    svc = HyperbandService()
    exp = _exp({"eta": 3, "r_l": 9, "resource_name": "epochs", "random_state": 1})
    req = api.GetSuggestionsRequest(experiment=exp, current_request_number=9)
    reply = svc.get_suggestions(req)
    pts = [{a.name: a.value for a in pa.assignments} for pa in reply.parameter_assignments]
    assert len(pts) == 9
    names = []
    for i, p in enumerate(pts):
        names.append(f"t{i}")
        add_trial(req, f"t{i}", p, loss=float(i))            # t0, t1, t2 are the best three
    del req.experiment.spec.algorithm.algorithm_settings[:]
    for s in reply.algorithm.algorithm_settings:
        x = req.experiment.spec.algorithm.algorithm_settings.add()
        x.name, x.value = s.name, (",".join(names) if s.name == "bracket_trials" else s.value)
    nxt = svc.get_suggestions(req)
    got = [{a.name: a.value for a in pa.assignments} for pa in nxt.parameter_assignments]
    assert len(got) == 3
    for g, p in zip(got, pts[:3]):
        assert g["epochs"] == "3" and all(g[k] == p[k] for k in p if k != "epochs")


@pytest.mark.parametrize("settings,frag", [({"eta": 3}, "r_l"), ({"eta": 1, "r_l": 9, "resource_name": "epochs"}, "eta"),
                                            ({"eta": 3, "r_l": 9}, "resource_name"), ({"eta": 3, "r_l": 9, "resource_name": "nope"}, "search space")])
def test_hyperband_validation(settings, frag):
    server, port = serve(HyperbandService(), port=0, host="127.0.0.1")
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    with pytest.raises(grpc.RpcError) as ei:
        SuggestionStub(ch).ValidateAlgorithmSettings(api.ValidateAlgorithmSettingsRequest(experiment=_exp(settings)))
    assert ei.value.code() == grpc.StatusCode.INVALID_ARGUMENT and frag in ei.value.details()
    ch.close()
    server.stop(0)
