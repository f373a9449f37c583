"""``api.v1.beta1.Suggestion`` servicers.

* ``SkoptService``  — algorithm ``bayesianoptimization`` (kubeflow/katib pkg/suggestion/v1beta1/skopt/service.py):
  GetSuggestions → BaseSkoptService → kubeflow_b200.Optimizer → libkbo (GPU).  ValidateAlgorithmSettings accepts the
  upstream settings (base_estimator, n_initial_points, acq_func, acq_optimizer, random_state) plus engine knobs as
  EXTRA settings, so an unchanged Experiment works and an extended one can opt in (SURVEY.md §5 "config").
* ``RandomService`` — algorithm ``random`` (upstream: the hyperopt service's random search): BASELINE.json config 1,
  gRPC plumbing on CPU, no GPU involved.
Errors: bad settings → INVALID_ARGUMENT; anything else → INTERNAL.
"""
from __future__ import annotations

import logging
import threading

import grpc
import numpy as np

from . import api_pb as api
from .internal import (AlgorithmSettingsError, HyperParameterSearchSpace, Trial, parse_settings)

logger = logging.getLogger(__name__)

SKOPT_SETTINGS = ("base_estimator", "n_initial_points", "acq_func", "acq_optimizer", "random_state")
ENGINE_SETTINGS = {"n_points": int, "kernel": str, "length_scale": float, "amplitude": float, "noise": float, "xi": float,
                   "kappa": float, "var_mode": str, "theta_grid": int, "theta_search": int, "theta_fit": str, "theta_fit_maxiter": int, "ard": int, "device": int, "candidate_backend": str}


def validate_skopt_settings(settings: dict) -> dict:
    """Returns the parsed keyword arguments or raises AlgorithmSettingsError (mirrors upstream OptimizerConfiguration)."""
    out = {}
    for name, value in settings.items():
        try:
            if name == "base_estimator":
                if value not in ("GP", "RF", "ET", "GBRT"):
                    raise AlgorithmSettingsError(f"base_estimator {value} is not supported in Bayesian optimization")
                if value != "GP":
                    raise AlgorithmSettingsError(f"base_estimator {value}: only GP runs on the GPU engine")
                out[name] = value
            elif name == "n_initial_points":
                if int(value) < 0:
                    raise AlgorithmSettingsError(f"n_initial_points should be great or equal than zero, got {value}")
                out[name] = int(value)
            elif name == "acq_func":
                if value not in ("gp_hedge", "LCB", "EI", "PI", "EIps", "PIps"):
                    raise AlgorithmSettingsError(f"acq_func {value} is not supported in Bayesian optimization")
                out[name] = value
            elif name == "acq_optimizer":
                if value not in ("auto", "sampling", "lbfgs"):
                    raise AlgorithmSettingsError(f"acq_optimizer {value} is not supported in Bayesian optimization")
                out[name] = value
            elif name == "random_state":
                if int(value) < 0:
                    raise AlgorithmSettingsError(f"random_state should be great or equal than zero, got {value}")
                out[name] = int(value)
            elif name in ENGINE_SETTINGS:
                v = ENGINE_SETTINGS[name](value)
                if name == "kernel" and v not in ("rbf", "matern52"):
                    raise AlgorithmSettingsError(f"kernel {v} must be rbf or matern52")
                if name == "var_mode" and v not in ("tc", "f64", "auto"):
                    raise AlgorithmSettingsError(f"var_mode {v} must be tc, f64 or auto")
                if name == "candidate_backend" and v not in ("torch", "numpy"):
                    raise AlgorithmSettingsError(f"candidate_backend {v} must be torch or numpy")
                if name == "theta_fit" and v != "lbfgs":
                    raise AlgorithmSettingsError(f"theta_fit {v} must be lbfgs")
                if name == "theta_search" and v < 0:
                    raise AlgorithmSettingsError(f"theta_search must be >= 0, got {value}")
                if name in ("n_points", "theta_grid") and v < 1:
                    raise AlgorithmSettingsError(f"{name} must be >= 1, got {value}")
                if name in ("length_scale", "amplitude") and not v > 0:
                    raise AlgorithmSettingsError(f"{name} must be > 0, got {value}")
                if name == "noise" and v < 0:
                    raise AlgorithmSettingsError(f"noise must be >= 0, got {value}")
                out[name] = v
            else:
                raise AlgorithmSettingsError(f"unknown setting {name} for algorithm bayesianoptimization")
        except AlgorithmSettingsError:
            raise
        except (TypeError, ValueError):
            raise AlgorithmSettingsError(f"failed to convert {value!r} for setting {name}")
    return out


def _reply_from(assignment_lists) -> "api.GetSuggestionsReply":
    reply = api.GetSuggestionsReply()
    for assignments in assignment_lists:
        pa = reply.parameter_assignments.add()
        for a in assignments:
            x = pa.assignments.add()
            x.name, x.value = a.name, a.value
    return reply


class _Base:
    algorithm_names: tuple = ()

    def _abort(self, context, code, msg):
        logger.warning("%s: %s", code, msg)
        context.set_code(code)
        context.set_details(msg)

    def ValidateAlgorithmSettings(self, request, context):
        try:
            self.validate(request.experiment)
        except AlgorithmSettingsError as e:
            self._abort(context, grpc.StatusCode.INVALID_ARGUMENT, str(e))
        except Exception as e:  # noqa: BLE001
            self._abort(context, grpc.StatusCode.INTERNAL, f"{type(e).__name__}: {e}")
        return api.ValidateAlgorithmSettingsReply()

    def GetSuggestions(self, request, context):
        try:
            return self.get_suggestions(request)
        except (AlgorithmSettingsError,) as e:
            self._abort(context, grpc.StatusCode.INVALID_ARGUMENT, str(e))
        except ValueError as e:
            self._abort(context, grpc.StatusCode.INVALID_ARGUMENT, str(e))
        except Exception as e:  # noqa: BLE001  (CUDA errors, not-PD, ...)
            logger.exception("GetSuggestions failed")
            self._abort(context, grpc.StatusCode.INTERNAL, f"{type(e).__name__}: {e}")
        return api.GetSuggestionsReply()


class SkoptService(_Base):
    algorithm_names = ("bayesianoptimization",)

    def __init__(self, engine_defaults: dict | None = None, max_experiments: int = 8):
        self._lock = threading.Lock()       # handlers may run concurrently (grpc thread pool): GPU work is serialised here
        self._services = {}                 # experiment name -> BaseSkoptService, least-recently-used first
        self.engine_defaults = dict(engine_defaults or {})
        self.last_ingest = None             # "scan" | "messages": how the last request's trials reached the optimizer
        self.max_experiments = int(max_experiments)   # each experiment owns a libkbo handle (K, W, scratch): bound the device memory

    def validate(self, experiment):
        name = experiment.spec.algorithm.algorithm_name
        if name != "bayesianoptimization":
            raise AlgorithmSettingsError(f"unknown algorithm name {name}")
        HyperParameterSearchSpace.convert(experiment)
        validate_skopt_settings(parse_settings(experiment))

    def get_suggestions(self, request):
        exp = request.experiment
        if exp.spec.algorithm.algorithm_name not in ("", "bayesianoptimization"):
            raise AlgorithmSettingsError(f"unknown algorithm name {exp.spec.algorithm.algorithm_name}")
        search_space = HyperParameterSearchSpace.convert(exp)
        settings = validate_skopt_settings(parse_settings(exp))
        from .ingest import LazyRequest
        lazy = request if isinstance(request, LazyRequest) and request.scanned else None
        # an experiment recreated under the same name, or whose search space / objective / settings changed, must not meet the
        # old optimizer (old dimensions, old history, old settings): the cache entry is valid for one fingerprint only
        fingerprint = (search_space.goal, exp.spec.objective.objective_metric_name,
                       tuple((p.name, p.type, p.min, p.max, tuple(p.list), p.step) for p in search_space.params),
                       tuple(sorted((k, repr(v)) for k, v in settings.items())))
        with self._lock:
            svc = self._services.get(exp.name)
            if svc is not None and svc.fingerprint != fingerprint:
                eng = getattr(self._services.pop(exp.name).skopt_optimizer, "_engine", None)
                if eng is not None:
                    eng.close()
                svc = None
            if svc is None:
                kw = dict(self.engine_defaults)
                kw.update(settings)
                from .base_service import BaseSkoptService
                svc = BaseSkoptService(search_space=search_space, **kw)
                svc.fingerprint = fingerprint
                self._services[exp.name] = svc
                while len(self._services) > self.max_experiments:       # evict the least recently used experiment's engine
                    old_name = next(iter(self._services))
                    old = self._services.pop(old_name)
                    eng = getattr(old.skopt_optimizer, "_engine", None)
                    if eng is not None:
                        eng.close()
            else:
                self._services[exp.name] = self._services.pop(exp.name)  # mark as most recently used
            # the wire scan tells the new trials from flat arrays; it declines (False) whenever a trial needs the message walk
            self.last_ingest = "scan" if (lazy is not None and svc.ingest(lazy)) else "messages"
            trials = [] if self.last_ingest == "scan" else Trial.convert(request.trials, skip_names=svc.told_trials)
            lists = svc.getSuggestions(trials, max(int(request.current_request_number), 0))
        return _reply_from(lists)


class RandomService(_Base):
    """Uniform random search over the feasible space (Katib algorithm ``random``); CPU only."""
    algorithm_names = ("random",)

    def __init__(self):
        self._rngs = {}
        self._lock = threading.Lock()

    def validate(self, experiment):
        if experiment.spec.algorithm.algorithm_name != "random":
            raise AlgorithmSettingsError(f"unknown algorithm name {experiment.spec.algorithm.algorithm_name}")
        HyperParameterSearchSpace.convert(experiment)
        for k, v in parse_settings(experiment).items():
            if k != "random_state":
                raise AlgorithmSettingsError(f"unknown setting {k} for algorithm random")
            try:
                if int(v) < 0:
                    raise ValueError
            except ValueError:
                raise AlgorithmSettingsError(f"random_state should be great or equal than zero, got {v}")

    def get_suggestions(self, request):
        from ..space import Categorical, Integer, Real, Space
        from .base_service import BaseSkoptService
        from .internal import CATEGORICAL, DISCRETE, DOUBLE, INTEGER
        exp = request.experiment
        ss = HyperParameterSearchSpace.convert(exp)
        st = parse_settings(exp)
        with self._lock:
            rng = self._rngs.get(exp.name)
            if rng is None:
                rng = self._rngs[exp.name] = np.random.default_rng(int(st["random_state"]) if "random_state" in st else None)
            dims = []
            for p in ss.params:
                dims.append(Integer(int(p.min), int(p.max), p.name) if p.type == INTEGER else
                            Real(float(p.min), float(p.max), p.name) if p.type == DOUBLE else Categorical(list(p.list), p.name))
            space = Space(dims)
            n = max(int(request.current_request_number), 0)
            pts = space.inverse_transform(space.rvs_transformed(n, rng, np.float64)) if n else []
        return _reply_from([BaseSkoptService.convert(ss, p) for p in pts])


class SobolService(_Base):
    """Quasi-random search over the feasible space (Katib algorithm ``sobol``, the second sampler of the goptuna service;
    SURVEY.md §2 row 18).  Scrambled Sobol points from SciPy's QMC engine, mapped through the same space transform as every
    other algorithm here; stateless across calls by skipping the points already handed out (= number of trials seen +
    number requested so far).  CPU only — there is no arithmetic to accelerate."""
    algorithm_names = ("sobol",)

    def __init__(self):
        self._lock = threading.Lock()
        self._issued = {}

    def validate(self, experiment):
        if experiment.spec.algorithm.algorithm_name != "sobol":
            raise AlgorithmSettingsError(f"unknown algorithm name {experiment.spec.algorithm.algorithm_name}")
        ss = HyperParameterSearchSpace.convert(experiment)
        for p in ss.params:
            if p.type in ("categorical", "discrete") and len(p.list) < 1:
                raise AlgorithmSettingsError(f"parameter {p.name!r}: empty feasible list")
        for k, v in parse_settings(experiment).items():
            if k != "random_state":
                raise AlgorithmSettingsError(f"unknown setting {k} for algorithm sobol")
            try:
                if int(v) < 0:
                    raise ValueError
            except ValueError:
                raise AlgorithmSettingsError(f"random_state should be great or equal than zero, got {v}")

    def get_suggestions(self, request):
        from scipy.stats import qmc
        from ..space import Categorical, Integer, Real, Space
        from .base_service import BaseSkoptService
        from .internal import DOUBLE, INTEGER
        exp = request.experiment
        ss = HyperParameterSearchSpace.convert(exp)
        st = parse_settings(exp)
        n = max(int(request.current_request_number), 0)
        with self._lock:
            start = self._issued.get(exp.name, 0)
            self._issued[exp.name] = start + n
        if n == 0:
            return _reply_from([])
        eng = qmc.Sobol(d=len(ss.params), scramble=True, seed=int(st.get("random_state", 0)))
        if start:
            eng.fast_forward(start)
        U = eng.random(n)
        dims = [Integer(int(p.min), int(p.max), p.name) if p.type == INTEGER else Real(float(p.min), float(p.max), p.name)
                if p.type == DOUBLE else Categorical(list(p.list), p.name) for p in ss.params]
        pts = []
        for u in U:
            pt = []
            for d, v in zip(dims, u):
                if isinstance(d, Categorical):
                    pt.append(d.categories[min(int(v * len(d.categories)), len(d.categories) - 1)])
                elif isinstance(d, Integer):
                    pt.append(min(d.low + int(v * (d.high - d.low + 1)), d.high))
                else:
                    pt.append(d.low + float(v) * (d.high - d.low))
            pts.append(pt)
        return _reply_from([BaseSkoptService.convert(ss, p) for p in pts])


class DispatchService(_Base):
    """One endpoint for several algorithms (upstream runs one Deployment per algorithm image; this routes by name)."""

    def __init__(self, services):
        self._by_name = {n: s for s in services for n in s.algorithm_names}

    def _pick(self, experiment):
        name = experiment.spec.algorithm.algorithm_name
        if name not in self._by_name:
            raise AlgorithmSettingsError(f"unknown algorithm name {name!r}; served: {sorted(self._by_name)}")
        return self._by_name[name]

    def validate(self, experiment):
        self._pick(experiment).validate(experiment)

    def get_suggestions(self, request):
        return self._pick(request.experiment).get_suggestions(request)
