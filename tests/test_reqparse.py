"""Request ingestion (kbo_req_* in include/kbo.h, suggestion/ingest.py) against protobuf's own parse of the same bytes and
against the message walk it replaces (internal.Trial.convert + BaseSkoptService.getSuggestions).  Host logic only: no GPU."""
import re

import numpy as np
import pytest
from google.protobuf.message import DecodeError

from kubeflow_b200.suggestion import api_pb as api
from kubeflow_b200.suggestion.ingest import MISSING, LazyRequest, hash64
from kubeflow_b200.suggestion.internal import HyperParameterSearchSpace, Trial

FLOAT_RE = re.compile(r"[+-]?(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?\Z", re.ASCII)   # ASCII digits only: float("１") is Python's business
INT_RE = re.compile(r"[+-]?\d{1,15}\Z", re.ASCII)
ODD = [" 1.0", "1_0", "inf", "nan", "0x10", "1e5", "-.5", "5.", ".", "", "+", "1e", "--1", "1.2.3", "１", "007", "+3", "-0", "1E-3", "abc",
       "123456789012345", "1234567890123456", "9" * 70, "0." + "3" * 80]


def _experiment(specs, objective=api.MINIMIZE, name="exp"):
    e = api.Experiment()
    e.name = name
    e.spec.objective.type = objective
    e.spec.objective.objective_metric_name = "loss"
    e.spec.algorithm.algorithm_name = "bayesianoptimization"
    for pname, ptype, lo, hi, lst in specs:
        p = e.spec.parameter_specs.parameters.add()
        p.name, p.parameter_type = pname, ptype
        if lst:
            p.feasible_space.list.extend(lst)
        else:
            p.feasible_space.min, p.feasible_space.max = str(lo), str(hi)
    return e


SPECS = [("lr", api.DOUBLE, 0.001, 0.5, None), ("layers", api.INT, 1, 9, None), ("opt", api.CATEGORICAL, 0, 0, ["sgd", "adam", "ftrl"]),
         ("momentum", api.DOUBLE, -1.0, 1.0, None), ("flip", api.CATEGORICAL, 0, 0, ["yes", "no"]), ("bs", api.DISCRETE, 0, 0, ["16", "32", "64"])]


def _random_request(rng, n_trials, weird=0.0, current=3):
    req = api.GetSuggestionsRequest(experiment=_experiment(SPECS), current_request_number=current, total_request_number=current + 4)
    pnames = [s[0] for s in SPECS]
    for i in range(n_trials):
        t = req.trials.add()
        t.name = f"trial-{rng.integers(0, max(2, n_trials // 2)) if rng.random() < 0.1 else i}"
        if rng.random() < 0.9:
            t.spec.objective.objective_metric_name = "loss"
        order = list(range(len(pnames)))
        if rng.random() < 0.3:
            rng.shuffle(order)
        for j in order:
            if rng.random() < 0.03:
                continue                                     # missing assignment
            reps = 2 if rng.random() < 0.05 else 1             # duplicate name: last wins
            for _ in range(reps):
                a = t.spec.parameter_assignments.assignments.add()
                a.name = pnames[j]
                kind = SPECS[j][1]
                if rng.random() < weird:
                    a.value = ODD[rng.integers(len(ODD))]
                elif kind == api.DOUBLE:
                    a.value = repr(float(rng.uniform(SPECS[j][2], SPECS[j][3])))
                elif kind == api.INT:
                    a.value = str(int(rng.integers(1, 10)))
                else:
                    a.value = str(rng.choice(SPECS[j][4]))
        if rng.random() < 0.1:
            a = t.spec.parameter_assignments.assignments.add()
            a.name, a.value = "not-a-parameter", "1"
        t.status.condition = int(rng.choice([api.SUCCEEDED, api.SUCCEEDED, api.SUCCEEDED, api.EARLYSTOPPED, 4, 1, 3]))
        t.status.start_time = "2026-01-01T00:00:00Z"
        for mname in rng.permutation(["loss", "accuracy", "loss"] if rng.random() < 0.1 else ["accuracy", "loss"]):
            if rng.random() < 0.05:
                continue
            m = t.status.observation.metrics.add()
            m.name = str(mname)
            m.value = ODD[rng.integers(len(ODD))] if rng.random() < weird else repr(float(rng.normal()))
    return req


def _reference_rows(req, pnames):
    """What Trial.convert + the by-name lookup of getSuggestions see, per trial of the request (None = filtered out)."""
    out = []
    for t in req.trials:
        conv = Trial.convert([t])
        if not conv:
            out.append(None)
            continue
        by = {a.name: a.value for a in conv[0].assignments}
        out.append((conv[0].name, conv[0].target_metric.value, [by.get(p) for p in pnames]))
    return out


@pytest.mark.parametrize("seed,weird", [(0, 0.0), (1, 0.0), (2, 0.3), (3, 0.3), (4, 1.0)])
def test_scan_matches_protobuf_and_trial_convert(seed, weird):
    rng = np.random.default_rng(seed)
    req = _random_request(rng, 300, weird)
    data = req.SerializeToString()
    lazy = LazyRequest.FromString(data)
    assert lazy.scanned and lazy.n_trials == 300
    assert lazy.current_request_number == req.current_request_number and lazy.total_request_number == req.total_request_number
    assert lazy.experiment == req.experiment and lazy.experiment.name == "exp"
    pnames = [s[0] for s in SPECS]
    tab = lazy.trial_table(pnames)
    ref = _reference_rows(req, pnames)
    for i, (t, r) in enumerate(zip(req.trials, ref)):
        assert lazy.text(tab.name_off[i], tab.name_len[i]) == t.name
        assert int(tab.name_hash[i]) == hash64(t.name)
        assert tab.condition[i] == t.status.condition
        assert bool(tab.usable[i]) == (r is not None)
        if r is None:
            continue
        _, target, vals = r
        assert lazy.text(tab.objective_off[i], tab.objective_len[i]) == target
        assert bool(tab.objective_flags[i] & 1) == bool(FLOAT_RE.match(target))
        if tab.objective_flags[i] & 1:
            assert tab.objective[i] == float(target)
        else:
            assert np.isnan(tab.objective[i])
        for j, v in enumerate(vals):
            if v is None:
                assert tab.value_len[i, j] == MISSING
                continue
            assert lazy.text(tab.value_off[i, j], tab.value_len[i, j]) == v
            is_f, is_i = bool(FLOAT_RE.match(v)), bool(INT_RE.match(v))
            assert int(tab.value_flags[i, j]) == (1 if is_f else 0) + (2 if is_i else 0), v
            if is_f:
                assert tab.values[i, j] == float(v), v          # strtod and Python's float() are both correctly rounded
            else:
                assert np.isnan(tab.values[i, j])
    # the message view is still there for the services that walk it
    assert len(lazy.trials) == 300 and lazy.trials[7].name == req.trials[7].name and lazy.ByteSize() == req.ByteSize()


def test_scan_edge_cases_and_malformed_bytes():
    empty = LazyRequest.FromString(b"")
    assert empty.scanned and empty.n_trials == 0 and empty.current_request_number == 0 and empty.experiment.name == ""
    assert empty.trial_table(["a"]).n == 0
    req = _random_request(np.random.default_rng(5), 20)
    data = req.SerializeToString()
    tab0 = LazyRequest.FromString(data).trial_table([])          # no parameters asked for: names / objective only
    assert tab0.values.shape == (20, 0) and tab0.usable.sum() > 0
    pn = [s_[0] for s_ in SPECS]
    full = LazyRequest.FromString(data).trial_table(pn)
    mask = np.zeros(20, dtype=bool); mask[[2, 5, 19]] = True
    part = LazyRequest.FromString(data).trial_table(pn, select=mask)   # numbers parsed for the selected trials only
    np.testing.assert_array_equal(part.name_hash, full.name_hash)
    np.testing.assert_array_equal(part.usable, full.usable)
    np.testing.assert_array_equal(part.value_len[mask], full.value_len[mask])
    np.testing.assert_array_equal(part.values[mask], full.values[mask])
    assert (part.value_len[~mask] == MISSING).all() and np.isnan(part.values[~mask]).all()
    with pytest.raises(ValueError):
        LazyRequest.FromString(data).trial_table(pn, select=mask[:5])
    # unknown fields of any wire type are skipped the way protobuf skips them
    extra = data + bytes([0x78, 0x05]) + bytes([0x81, 0x01]) + b"\x00" * 8 + bytes([0x8d, 0x01]) + b"\x00" * 4 + bytes([0x92, 0x01, 0x03]) + b"abc"
    lz = LazyRequest.FromString(extra)
    assert lz.scanned and lz.n_trials == 20 and api.GetSuggestionsRequest.FromString(extra).trials[3].name == lz.trials[3].name
    for cut in (1, 7, len(data) // 2, len(data) - 1):            # truncated: the scan declines, protobuf raises as it always did
        with pytest.raises(DecodeError):
            LazyRequest.FromString(data[:cut]).trials


class _NoAsk:
    def __init__(self, opt):
        self.opt = opt

    def __getattr__(self, k):
        return getattr(self.opt, k)

    def ask(self, n_points=None):
        return []


def _service(objective):
    from kubeflow_b200.suggestion.base_service import BaseSkoptService
    space = HyperParameterSearchSpace.convert(_experiment(SPECS, objective))
    svc = BaseSkoptService(search_space=space, n_initial_points=3, random_state=0)
    svc.skopt_optimizer = _NoAsk(svc.skopt_optimizer)          # tell() is host-only; ask() needs the GPU
    return svc


@pytest.mark.parametrize("objective", [api.MINIMIZE, api.MAXIMIZE])
def test_ingest_tells_the_optimizer_exactly_what_the_message_walk_tells(objective):
    rng = np.random.default_rng(11)
    req = _random_request(rng, 400)
    # keep only trials the message walk accepts (complete assignments); the declined cases are tested below
    pnames = [s[0] for s in SPECS]
    good = api.GetSuggestionsRequest(experiment=_experiment(SPECS, objective), current_request_number=2)
    for t, r in zip(req.trials, _reference_rows(req, pnames)):
        if r is None or all(v is not None for v in r[2]):
            good.trials.add().CopyFrom(t)
    slow, fast = _service(objective), _service(objective)
    half = api.GetSuggestionsRequest(experiment=good.experiment, current_request_number=1)
    for t in good.trials[:150]:
        half.trials.add().CopyFrom(t)
    for r in (half, good, good):                                 # second call tells only the new trials, third tells nothing
        slow.getSuggestions(Trial.convert(r.trials, skip_names=slow.told_trials), 1)
        assert fast.ingest(LazyRequest.FromString(r.SerializeToString())) is True
        assert fast.told_trials == slow.told_trials and len(slow.told_trials) > 0
        assert fast.skopt_optimizer.Xi == slow.skopt_optimizer.Xi
        assert [type(v) for v in fast.skopt_optimizer.Xi[-1]] == [type(v) for v in slow.skopt_optimizer.Xi[-1]]
        assert fast.skopt_optimizer.yi == slow.skopt_optimizer.yi
        np.testing.assert_array_equal(fast.skopt_optimizer._Xt, slow.skopt_optimizer._Xt)
    # mixing the paths on one service: names told through the messages are recognised by the scan
    slow.ingest(LazyRequest.FromString(good.SerializeToString()))
    assert len(slow.skopt_optimizer.yi) == len(fast.skopt_optimizer.yi)


@pytest.mark.parametrize("field,value", [("lr", " 0.1"), ("lr", "1_0"), ("lr", "nan"), ("layers", "3.0"), ("layers", "x"), ("loss", "inf"),
                                         ("lr", None)])
def test_ingest_declines_what_needs_the_reference_conversion(field, value):
    """Values Python's float()/int() may accept or reject in their own way are left to the message walk, untouched."""
    exp = _experiment(SPECS)
    req = api.GetSuggestionsRequest(experiment=exp, current_request_number=1)
    for i in range(3):
        t = req.trials.add()
        t.name = f"t{i}"
        t.spec.objective.objective_metric_name = "loss"
        vals = {"lr": "0.1", "layers": "3", "opt": "adam", "momentum": "0.5", "flip": "no", "bs": "32"}
        if i == 1 and field != "loss":
            if value is None:
                del vals[field]
            else:
                vals[field] = value
        for k, v in vals.items():
            a = t.spec.parameter_assignments.assignments.add()
            a.name, a.value = k, v
        t.status.condition = api.SUCCEEDED
        m = t.status.observation.metrics.add()
        m.name, m.value = "loss", (value if (i == 1 and field == "loss") else "0.25")
    svc = _service(api.MINIMIZE)
    assert svc.ingest(LazyRequest.FromString(req.SerializeToString())) is False
    assert svc.told_trials == set() and svc.skopt_optimizer.yi == []
    # and the message walk then behaves as the reference does: converts what Python converts, raises what Python raises
    walk = lambda: svc.getSuggestions(Trial.convert(req.trials), 1)   # noqa: E731
    if (field, value) in (("lr", " 0.1"), ("lr", "1_0"), ("lr", "nan"), ("loss", "inf")):
        walk()
        assert len(svc.skopt_optimizer.yi) == 3
    else:
        with pytest.raises(ValueError):
            walk()
