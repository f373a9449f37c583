"""``Optimizer`` — the host-side mirror of ``skopt.Optimizer`` as Katib's skopt service drives it
(upstream kubeflow/katib pkg/suggestion/v1beta1/skopt/base_service.py: ``tell(X, y)`` then ``ask()`` per
requested assignment).  Differences, stated once:

* θ is FIXED per request (length scales, amplitude, noise are settings), where skopt refits them by L-BFGS on the
  log-marginal likelihood ($SK/_gpr.py:299-339) — bit-reproducing that optimiser is outside the 1e-5 contract
  (SURVEY.md §7).  Two GPU-side substitutes, both maximising the device-computed LML (one FP64 fit per candidate θ):
  ``theta_grid`` = G > 1 tries G multiples of the length scale (geometric, 0.25×…4×); ``theta_search`` = K > 0 additionally
  draws K random (length-scale multiple, noise) pairs log-uniformly from [0.2, 5] × [1e-6, 1e-1] (seeded) — the gradient-free
  analogue of skopt's restarts; ``theta_fit="lbfgs"`` then polishes (log ℓ — per dimension with ``ard=True`` —, log noise)
  with SciPy's L-BFGS-B exactly as sklearn does ($SK/_gpr.py:658-667), every objective/gradient evaluation being one
  device fit + ``kbo_lml_grad``.
* ``acq_optimizer`` is "sampling" over ``n_points`` candidates (skopt default 10 000; here 65 536 by default and
  millions are cheap) — ``lbfgs`` polishing and ``gp_hedge`` are accepted by ValidateAlgorithmSettings and mapped to
  sampling / EI.
* ``ask(n_points=k)`` uses skopt's constant-liar "cl_min" strategy: k sequential asks with the lie y = min(y).
"""
from __future__ import annotations

import numpy as np

from .space import Space


class Optimizer:
    def __init__(self, dimensions, base_estimator="GP", n_initial_points=10, acq_func="EI", acq_optimizer="sampling",
                 random_state=None, *, n_points=65536, kernel="matern52", length_scale=None, amplitude=1.0, noise=1e-3, xi=0.01,
                 kappa=1.96, var_mode="auto", theta_grid=1, theta_search=0, theta_fit=None, theta_fit_maxiter=25, ard=False, device=0, engine=None,
                 candidate_backend="torch", incremental=True, shard=False):
        if str(base_estimator).upper() != "GP":
            raise ValueError("base_estimator must be GP (RF/ET/GBRT are not part of the GPU path)")
        self.space = dimensions if isinstance(dimensions, Space) else Space(dimensions)
        self.n_initial_points = int(n_initial_points)
        acq = str(acq_func)
        self.acq = {"EI": "ei", "LCB": "lcb", "PI": "pi", "gp_hedge": "ei", "EIps": "ei", "PIps": "pi"}.get(acq)
        if self.acq is None:
            raise ValueError(f"unknown acq_func {acq_func!r}")
        if acq_optimizer not in ("auto", "sampling", "lbfgs"):
            raise ValueError(f"unknown acq_optimizer {acq_optimizer!r}")
        self.rng = np.random.default_rng(random_state)
        self.n_points = int(n_points)
        self.kernel, self.amplitude, self.noise, self.xi, self.kappa = kernel, amplitude, noise, xi, kappa
        self.length_scale = length_scale
        self.var_mode, self.theta_grid, self.theta_search, self.device = var_mode, int(theta_grid), int(theta_search), device
        if theta_fit not in (None, "lbfgs"):
            raise ValueError("theta_fit must be None or 'lbfgs'")
        self.theta_fit, self.theta_fit_maxiter, self.ard = theta_fit, int(theta_fit_maxiter), bool(ard)
        self.last_theta = None
        if candidate_backend not in ("torch", "numpy"):
            raise ValueError("candidate_backend must be 'torch' (sampled on the device) or 'numpy'")
        self.candidate_backend = candidate_backend
        self._tgen = None
        self._crng = None
        self.last_local_best = None
        self._seed = random_state
        self.Xi, self.yi = [], []
        self._Xt = np.empty((0, self.space.transformed_n_dims))   # transformed rows, one per told point
        self._engine = engine
        self.incremental = bool(incremental)   # False: refit on every ask, as skopt does
        # Sharded sweep: every rank of the default process group holds the same optimizer (same tells, same
        # seed) and sweeps its own n_points/world candidates; one all-gather picks the winner, its owner broadcasts the row.
        # Opt-in (the SPMD server sets it): every rank must then make the same calls in the same order.
        self.shard = bool(shard) and self._world()[1] > 1
        if self.shard and random_state is None:
            import torch.distributed as dist
            box = [int(np.random.default_rng().integers(0, 2 ** 31))]
            dist.broadcast_object_list(box, src=0)          # ranks must draw the same initial points
            random_state = box[0]
            self._seed = random_state
            self.rng = np.random.default_rng(random_state)
        self._fit_state = None        # (θ key, rows in the engine, digest of those rows): lets a request that only adds trials append
        self.last_fit = None          # "fit" | "append" | "reuse": what the last ask did to the engine (diagnostics, tests)
        self.last_best = None

    # ------------------------------------------------------------------------------------------------------
    def tell(self, x, y, xt=None):
        """``xt``: the rows already in the transformed space (the request scan computes them column-wise); must equal
        ``space.transform(x)``."""
        if len(x) and not isinstance(x[0], (list, tuple)):
            x, y = [x], [y]
            xt = None if xt is None else np.atleast_2d(xt)
        if len(x) != len(y):
            raise ValueError("x and y must have the same length")
        rows = [list(p) for p in x]
        self.Xi.extend(rows)
        self.yi.extend([float(v) for v in y])
        if rows:
            new = self.space.transform(rows) if xt is None else np.asarray(xt, dtype=np.float64).reshape(len(rows), -1)
            self._Xt = np.concatenate([self._Xt, new])

    def _get_engine(self, ls):
        from .gp import GPEngine
        if self._engine is None:
            self._engine = GPEngine(self.device, kernel=self.kernel, length_scale=ls, amplitude=self.amplitude, noise=self.noise,
                                    acq=self.acq, xi=self.xi, kappa=self.kappa, var_mode=self.var_mode)
        e = self._engine
        e.kernel, e.acq, e.amplitude, e.noise, e.xi, e.kappa = self.kernel, self.acq, self.amplitude, self.noise, self.xi, self.kappa
        e.length_scale = np.atleast_1d(np.asarray(ls, dtype=np.float64))
        return e

    def _tell_engine(self, eng, Xt, ya, max_append=16):
        """Bring the engine to the history (Xt, ya) at its current θ.  Katib resends the whole history on every request and
        the constant liar extends it one row at a time, so most calls differ from what the engine already holds only past a
        long common prefix of rows: the engine keeps that prefix (kbo_fit_rebase — the leading blocks of L and L⁻¹ are the
        factors of the shorter history; targets are replaced wholesale, so a lie that became an observation costs nothing) and
        the few new rows are appended (kbo_fit_append, O(N²) each) while there is room in the 64-row pitch.  Anything else —
        a different θ, an edited early row, many new rows — refits (O(N³)) as skopt does on every tell."""
        key = (eng.kernel, tuple(np.atleast_1d(eng.length_scale).tolist()), eng.amplitude, eng.noise, eng.normalize_y, eng.var_mode,
               eng.acq, eng.xi, eng.kappa)
        n, st = len(ya), self._fit_state
        self._fit_state = None          # stays None if anything below raises: the engine state is then unknown
        done = None
        if self.incremental and st is not None and st[0] == key and n >= 1:
            Xh, yh = st[1], st[2]
            m = min(n, len(yh))
            diff = np.flatnonzero((Xt[:m] != Xh[:m]).any(axis=1))
            p = int(diff[0]) if len(diff) else m          # rows [0, p) of the engine are rows [0, p) of the new history
            if p >= 1 and n - p <= max_append and p + eng.room() + (len(yh) - p) >= n:
                if p == n == len(yh) and np.array_equal(ya, yh):
                    done = "reuse"
                else:
                    if p != len(yh) or not np.array_equal(ya[:p], yh[:p]):   # rows dropped or targets changed
                        eng.rebase(p, ya[:p])
                    for i in range(p, n):
                        eng.append(Xt[i], ya[i])
                    done = "append" if n > p else "rebase"
        if done is None:
            eng.tell(Xt, ya)
            done = "fit"
        self.last_fit = done
        self._fit_state = (key, np.array(Xt, dtype=np.float64, copy=True), np.array(ya, dtype=np.float64, copy=True))

    @staticmethod
    def _world():
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        return 0, 1

    def _sweep(self, eng, cand, global_offset=0):
        try:
            return eng.ask(cand, global_offset=global_offset)
        except Exception:
            self._fit_state = None      # e.g. an appended row made K singular: the next ask must refit from the history
            raise

    def _default_ls(self):
        return 0.3 * np.sqrt(self.space.transformed_n_dims) if self.length_scale is None else self.length_scale

    def _ask_one(self, X, y):
        if len(y) < max(self.n_initial_points, 1):
            return self.space.inverse_transform(self.space.rvs_transformed(1, self.rng, np.float64))[0]
        n_told = len(self._Xt)
        Xt = self._Xt
        if len(X) > n_told:          # constant-liar rows appended by ask(n_points=k)
            Xt = np.concatenate([Xt, self.space.transform(X[n_told:])])
        ya = np.asarray(y, dtype=np.float64)
        base = np.atleast_1d(np.asarray(self._default_ls(), dtype=np.float64))
        eng = self._get_engine(base)
        if self.theta_grid > 1 or self.theta_search > 0:   # choose θ by the GPU log-marginal likelihood
            cands = [(float(m), self.noise) for m in (np.geomspace(0.25, 4.0, self.theta_grid) if self.theta_grid > 1 else [1.0])]
            if self.theta_search > 0:
                tr = np.random.default_rng(0 if self._seed is None else int(self._seed) + 17)
                cands += [(float(np.exp(tr.uniform(np.log(0.2), np.log(5.0)))), float(np.exp(tr.uniform(np.log(1e-6), np.log(1e-1)))))
                          for _ in range(self.theta_search)]
            best_lml, best_t = -np.inf, (1.0, self.noise)
            if hasattr(eng, "lml_batch"):
                # all candidates in a few calls: the factorisations of a batch run concurrently on the device (kbo_lml_batch) —
                # the chain of diagonal blocks of one hides behind the others' — instead of one full fit per θ
                gb = max(1, min(8, int(2 ** 31 // max(1, 8 * len(ya) * len(ya)))))      # ≤ 2 GiB of Gram matrices per batch
                for i in range(0, len(cands), gb):
                    part = cands[i:i + gb]
                    lmls = eng.lml_batch(Xt, ya, [dict(length_scale=base * m, noise=nz) for m, nz in part])
                    for (m, nz), lml in zip(part, lmls):
                        if np.isfinite(lml) and lml > best_lml:
                            best_lml, best_t = float(lml), (m, nz)
            else:
                self._fit_state = None
                for m, nz in cands:
                    eng.length_scale, eng.noise = base * m, nz
                    eng.tell(Xt, ya)
                    try:
                        lml = eng.fit_info()["lml"]
                    except Exception:   # not positive definite at this θ
                        continue
                    if lml > best_lml:
                        best_lml, best_t = lml, (m, nz)
            eng.length_scale, eng.noise = base * best_t[0], best_t[1]
            self.last_theta = dict(length_scale=eng.length_scale.copy(), noise=eng.noise, lml=best_lml)
        if self.theta_fit == "lbfgs":
            from scipy.optimize import minimize
            Dt = self.space.transformed_n_dims
            ls0 = np.broadcast_to(np.atleast_1d(eng.length_scale), (Dt,)).copy() if self.ard else np.atleast_1d(eng.length_scale)[:1].copy()
            x0 = np.concatenate([np.log(ls0), [np.log(max(eng.noise, 1e-8))]])
            bounds = [(np.log(1e-2), np.log(1e2))] * len(ls0) + [(np.log(1e-8), np.log(1.0))]

            self._fit_state = None

            def negative_lml(t):
                eng.length_scale, eng.noise = np.exp(t[:-1]), float(np.exp(t[-1]))
                eng.tell(Xt, ya)
                try:
                    lml, g = eng.lml_grad()            # g: (log amp, log noise, log ℓ…)
                except Exception:                       # not positive definite at this θ
                    return 1e25, np.zeros_like(t)
                return -lml, -np.concatenate([g[2:], g[1:2]])

            res = minimize(negative_lml, x0, jac=True, method="L-BFGS-B", bounds=bounds, options=dict(maxiter=self.theta_fit_maxiter))
            eng.length_scale, eng.noise = np.exp(res.x[:-1]), float(np.exp(res.x[-1]))
            self.last_theta = dict(length_scale=eng.length_scale.copy(), noise=eng.noise, lml=-float(res.fun), nfev=int(res.nfev))
        self._tell_engine(eng, Xt, ya)
        rank, world = self._world() if self.shard else (0, 1)
        from .dist import global_argmax, shard_rows
        if world > 1 and self.n_points < world:     # identical on every rank, before any collective: all raise together
            raise ValueError(f"n_points={self.n_points} must be >= the number of ranks ({world}) when the grid is sharded")
        lo, hi = shard_rows(self.n_points, rank, world)
        if self.candidate_backend == "torch":
            import torch
            dev = torch.device("cuda", self.device)
            if self._tgen is None:
                self._tgen = torch.Generator(device=dev)
                base = int(self._seed) if self._seed is not None else int(self.rng.integers(0, 2 ** 31))
                self._tgen.manual_seed(base + 1000003 * rank)
            cand = self.space.rvs_transformed_torch(hi - lo, self._tgen, dev)
        else:
            if self._crng is None:
                self._crng = self.rng if world == 1 else np.random.default_rng([int(self._seed), rank])
            cand = self.space.rvs_transformed(hi - lo, self._crng, np.float32)
        failed = None
        try:
            best = self._sweep(eng, cand, lo)
        except Exception as e:  # noqa: BLE001
            if world == 1:
                raise
            best, failed = None, f"{type(e).__name__}: {e}"   # still join the exchange: every rank raises there, none hangs
        self.last_local_best = best
        li = best.index - lo if best is not None else 0
        if world > 1:
            import torch
            import torch.distributed as dist
            cdev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            best = global_argmax(best, device=cdev, failed=failed)
            owner = next(r for r in range(world) if shard_rows(self.n_points, r, world)[0] <= best.index < shard_rows(self.n_points, r, world)[1])
            rowt = torch.zeros(self.space.transformed_n_dims, dtype=torch.float64, device=cdev)
            if rank == owner:
                li = best.index - lo
                src = cand[li] if isinstance(cand, torch.Tensor) else torch.from_numpy(np.asarray(cand[li]))
                rowt.copy_(src.to(torch.float64))
            dist.broadcast(rowt, src=owner)
            row = rowt.cpu().numpy()[None, :]
        elif self.candidate_backend == "torch":
            import torch
            row = cand[li:li + 1].to(torch.float64).cpu().numpy()
        else:
            row = cand[li:li + 1].astype(np.float64)
        self.last_best = best
        return self.space.inverse_transform(row)[0]

    def ask(self, n_points=None):
        if n_points is None:
            return self._ask_one(self.Xi, self.yi)
        X, y, out = list(self.Xi), list(self.yi), []
        for _ in range(int(n_points)):
            x = self._ask_one(X, y)
            out.append(x)
            if y:                      # constant liar, "cl_min"
                X.append(x)
                y.append(min(y))
        return out
