"""``cmaes`` — Katib's goptuna-backed CMA-ES algorithm on the GPU sampler ``kubeflow_b200.cmaes``.

Search happens in the normalised cube [0,1]^D (reals and integers; categorical/discrete parameters are rejected, as
goptuna's CMA sampler does), mean0 = 0.5, sigma0 = 1/6 unless ``sigma`` is set; samples are clipped to the cube.
A generation's λ points are handed out across GetSuggestions calls; when all λ results are back the population is told
and the next generation is sampled.  Samples are tracked BY INDEX: two samples that round or clip to the same assignment
strings (integer parameters, cube boundary) are two pending entries under one key, matched first-in first-out; a sample
whose trial ended without an objective (FAILED / KILLED / METRICSUNAVAILABLE) goes back to the head of the queue and is
handed out again — goptuna likewise keeps sampling the running generation until popsize trials are COMPLETE — so neither
duplicates nor failures can stall the generation.  If the controller asks for more points than the generation has left,
the extra points are independent draws from the current distribution (they do not enter the update), which is what
goptuna's relative/independent sampling split amounts to.
"""
from __future__ import annotations

import threading

import numpy as np

from . import api_pb as api
from .internal import (CATEGORICAL, DISCRETE, DOUBLE, INTEGER, MAX_GOAL, AlgorithmSettingsError, Assignment,
                       HyperParameterSearchSpace, Trial, parse_settings)
from .service import _Base, _reply_from


def validate_cmaes_settings(settings: dict, n_params: int) -> dict:
    out = {}
    for k, v in settings.items():
        try:
            if k == "random_state":
                out[k] = int(v)
                if out[k] < 0:
                    raise AlgorithmSettingsError(f"random_state should be great or equal than zero, got {v}")
            elif k == "sigma":
                out[k] = float(v)
                if not out[k] > 0:
                    raise AlgorithmSettingsError(f"sigma should be positive, got {v}")
            elif k == "popsize":
                out[k] = int(v)
                if not 4 <= out[k] <= 8192:
                    raise AlgorithmSettingsError(f"popsize must be in [4, 8192], got {v}")
            elif k in ("restart_strategy", "device"):
                out[k] = v
            else:
                raise AlgorithmSettingsError(f"unknown setting {k} for algorithm cmaes")
        except ValueError:
            raise AlgorithmSettingsError(f"failed to convert {v!r} for setting {k}")
    if n_params < 2:
        raise AlgorithmSettingsError("cmaes only supports two or more dimensional continuous search space")   # goptuna's rule
    return out


class _Experiment:
    def __init__(self, space, D, settings):
        from ..cmaes import CmaEs
        self.es = CmaEs(np.full(D, 0.5), settings.get("sigma", 1.0 / 6.0), popsize=settings.get("popsize"),
                        seed=settings.get("random_state", 0), device=int(settings.get("device", 0)))
        self.queue = []          # [(sample index, assignments)] of the current generation not handed out yet
        self.assign = {}         # sample index -> assignments of the current generation
        self.pending = {}        # key -> [sample indices handed out, no result yet] (first in, first out)
        self.fitness = {}        # sample index -> value
        self.seen = set()        # trial names already accounted for (finished with or without an objective)
        self.extra_rng = np.random.default_rng(settings.get("random_state", 0) + 7919)


class CmaesService(_Base):
    algorithm_names = ("cmaes",)

    def __init__(self):
        self._lock = threading.Lock()
        self._exps = {}

    @staticmethod
    def _space(experiment):
        ss = HyperParameterSearchSpace.convert(experiment)
        for p in ss.params:
            if p.type in (CATEGORICAL, DISCRETE):
                raise AlgorithmSettingsError(f"cmaes does not support categorical/discrete parameter {p.name!r}")
        return ss

    def validate(self, experiment):
        if experiment.spec.algorithm.algorithm_name != "cmaes":
            raise AlgorithmSettingsError(f"unknown algorithm name {experiment.spec.algorithm.algorithm_name}")
        ss = self._space(experiment)
        validate_cmaes_settings(parse_settings(experiment), len(ss.params))

    @staticmethod
    def _to_assignments(ss, u):
        out = []
        for p, v in zip(ss.params, np.clip(u, 0.0, 1.0)):
            lo, hi = float(p.min), float(p.max)
            x = lo + float(v) * (hi - lo)
            out.append(Assignment(p.name, str(int(round(x))) if p.type == INTEGER else repr(x)))
        return out

    @staticmethod
    def _key(assignments):
        return tuple((a.name, a.value) for a in assignments)

    def _new_generation(self, st, ss):
        X = st.es.ask().cpu().numpy()
        st.queue = [(i, self._to_assignments(ss, X[i])) for i in range(X.shape[0])]
        st.assign = {i: a for i, a in st.queue}
        st.pending = {}
        st.fitness = {}

    def get_suggestions(self, request):
        exp = request.experiment
        ss = self._space(exp)
        settings = validate_cmaes_settings(parse_settings(exp), len(ss.params))
        trials = Trial.convert(request.trials)
        sign = -1.0 if ss.goal == MAX_GOAL else 1.0
        with self._lock:
            st = self._exps.get(exp.name)
            if st is None:
                st = self._exps[exp.name] = _Experiment(ss, len(ss.params), settings)
                self._new_generation(st, ss)
            for t in trials:
                if t.name in st.seen:
                    continue
                st.seen.add(t.name)
                waiting = st.pending.get(self._key(t.assignments))
                if waiting:
                    st.fitness[waiting.pop(0)] = sign * float(t.target_metric.value)
            # trials that ended without an objective: their sample is handed out again
            for t in request.trials:
                # (a SUCCEEDED / EARLYSTOPPED trial still unseen here carried no objective metric: same treatment)
                if t.name in st.seen or t.status.condition not in (api.FAILED, api.KILLED, api.METRICSUNAVAILABLE, api.SUCCEEDED,
                                                                   api.EARLYSTOPPED):
                    continue
                st.seen.add(t.name)
                waiting = st.pending.get(self._key(Assignment.convert(t.spec.parameter_assignments.assignments)))
                if waiting:
                    i = waiting.pop(0)
                    st.queue.insert(0, (i, st.assign[i]))
            if len(st.fitness) == st.es.popsize:
                st.es.tell(np.array([st.fitness[i] for i in range(st.es.popsize)]))
                self._new_generation(st, ss)
            lists = []
            for _ in range(max(int(request.current_request_number), 0)):
                if st.queue:
                    i, a = st.queue.pop(0)
                    st.pending.setdefault(self._key(a), []).append(i)
                    lists.append(a)
                else:   # generation exhausted but not finished: independent draw from the current distribution
                    s = st.es.state()
                    z = st.extra_rng.standard_normal(st.es.D)
                    lists.append(self._to_assignments(ss, s["mean"] + s["sigma"] * (s["B"] @ (s["d"] * z))))
        return _reply_from(lists)
