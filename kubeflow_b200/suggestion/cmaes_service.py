"""``cmaes`` — Katib's goptuna-backed CMA-ES algorithm on the GPU sampler ``kubeflow_b200.cmaes``.

Search happens in the normalised cube [0,1]^D (reals and integers; categorical/discrete parameters are rejected, as
goptuna's CMA sampler does), mean0 = 0.5, sigma0 = 1/6 unless ``sigma`` is set; samples are clipped to the cube.
A generation's λ points are handed out across GetSuggestions calls; when all λ results are back the population is told
and the next generation is sampled.  If the controller asks for more points than the generation has left, the extra
points are independent draws from the current distribution (they do not enter the update), which is what goptuna's
relative/independent sampling split amounts to.
"""
from __future__ import annotations

import threading

import numpy as np

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
        self.pending = {}        # key -> sample index
        self.fitness = {}        # sample index -> value
        self.seen = set()
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
        st.pending = {self._key(a): i for i, a in st.queue}
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
                i = st.pending.get(self._key(t.assignments))
                if i is not None and i not in st.fitness:
                    st.fitness[i] = sign * float(t.target_metric.value)
            if len(st.fitness) == st.es.popsize:
                st.es.tell(np.array([st.fitness[i] for i in range(st.es.popsize)]))
                self._new_generation(st, ss)
            lists = []
            for _ in range(max(int(request.current_request_number), 0)):
                if st.queue:
                    lists.append(st.queue.pop(0)[1])
                else:   # generation exhausted but not finished: independent draw from the current distribution
                    s = st.es.state()
                    z = st.extra_rng.standard_normal(st.es.D)
                    lists.append(self._to_assignments(ss, s["mean"] + s["sigma"] * (s["B"] @ (s["d"] * z))))
        return _reply_from(lists)
