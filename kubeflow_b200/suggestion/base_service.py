"""``BaseSkoptService`` — mirror of kubeflow/katib pkg/suggestion/v1beta1/skopt/base_service.py: holds the
optimizer for one experiment, tells it the trials it has not seen yet and asks for ``current_request_number``
new assignments."""
from __future__ import annotations

import logging

from ..optimizer import Optimizer
from ..space import Categorical, Integer, Real
from .internal import CATEGORICAL, DISCRETE, DOUBLE, INTEGER, MAX_GOAL, Assignment

logger = logging.getLogger(__name__)


class BaseSkoptService:
    def __init__(self, base_estimator="GP", n_initial_points=10, acq_func="gp_hedge", acq_optimizer="auto", random_state=None,
                 search_space=None, **engine_settings):
        self.base_estimator, self.n_initial_points = base_estimator, n_initial_points
        self.acq_func, self.acq_optimizer, self.random_state = acq_func, acq_optimizer, random_state
        self.search_space = search_space
        self.engine_settings = engine_settings
        self.skopt_optimizer = None
        self.told_trials = set()
        self.create_optimizer()

    def create_optimizer(self):
        dims = []
        for p in self.search_space.params:
            if p.type == INTEGER:
                dims.append(Integer(int(p.min), int(p.max), name=p.name))
            elif p.type == DOUBLE:
                dims.append(Real(float(p.min), float(p.max), name=p.name))
            elif p.type in (CATEGORICAL, DISCRETE):
                dims.append(Categorical(list(p.list), name=p.name))
        self.skopt_optimizer = Optimizer(dims, base_estimator=self.base_estimator, n_initial_points=self.n_initial_points,
                                         acq_func=self.acq_func, acq_optimizer=self.acq_optimizer, random_state=self.random_state,
                                         **self.engine_settings)

    def getSuggestions(self, trials, current_request_number):
        """trials: internal.Trial list (all completed trials, resent every call); returns a list of Assignment lists."""
        skopt_suggested, loss_for_skopt = [], []
        for trial in trials:
            if trial.name in self.told_trials:
                continue
            row = []
            by_name = {a.name: a.value for a in trial.assignments}
            for param in self.search_space.params:
                value = by_name.get(param.name)
                if value is None:
                    raise ValueError(f"trial {trial.name!r} has no assignment for parameter {param.name!r}")
                row.append(int(value) if param.type == INTEGER else float(value) if param.type == DOUBLE else value)
            loss = float(trial.target_metric.value)
            if self.search_space.goal == MAX_GOAL:
                loss = -1.0 * loss
            skopt_suggested.append(row)
            loss_for_skopt.append(loss)
            self.told_trials.add(trial.name)
        if skopt_suggested:
            self.skopt_optimizer.tell(skopt_suggested, loss_for_skopt)
        points = self.skopt_optimizer.ask(n_points=current_request_number)
        return [self.convert(self.search_space, p) for p in points]

    @staticmethod
    def convert(search_space, skopt_suggested):
        out = []
        for param, v in zip(search_space.params, skopt_suggested):
            if param.type == INTEGER:
                out.append(Assignment(param.name, str(int(v))))
            elif param.type == DOUBLE:
                out.append(Assignment(param.name, repr(float(v))))
            else:
                out.append(Assignment(param.name, str(v)))
        return out
