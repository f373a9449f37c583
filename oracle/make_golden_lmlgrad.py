"""Golden LML gradients from the REAL scikit-learn GPR (ConstantKernel·{RBF,Matern} + WhiteKernel, alpha = 0,
``log_marginal_likelihood(theta, eval_gradient=True)``, $SK/_gpr.py:541-655) -> tests/golden/lmlgrad_cases.npz.
θ order here: (log amplitude, log noise, log ℓ_1..P).  Run:  python -m oracle.make_golden_lmlgrad"""
import os

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel


def main():
    r = np.random.default_rng(77)
    out = {}
    for name, N, D, kind, ard in (("matern_iso", 40, 3, "matern52", False), ("rbf_ard", 35, 4, "rbf", True),
                                  ("matern_ard", 50, 5, "matern52", True), ("rbf_iso", 30, 2, "rbf", False)):
        X = r.random((N, D))
        y = np.sin(3 * X.sum(1) / np.sqrt(D)) + 0.1 * r.standard_normal(N)
        ls = (0.3 * np.sqrt(D) * r.uniform(0.7, 1.4, D)) if ard else 0.3 * np.sqrt(D)
        amp, noise = 1.7, 3e-3
        base = RBF(length_scale=ls) if kind == "rbf" else Matern(length_scale=ls, nu=2.5)
        gpr = GaussianProcessRegressor(kernel=ConstantKernel(amp) * base + WhiteKernel(noise), alpha=0.0, normalize_y=True,
                                       optimizer=None).fit(X, y)
        lml, g = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
        g_ours = np.concatenate([[g[0], g[-1]], g[1:-1]])      # sklearn: [log amp, log ℓ…, log noise]
        out[name] = dict(X=X, y=y, ls=np.atleast_1d(ls), amp=amp, noise=noise, kind=kind, lml=lml, grad=g_ours)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lmlgrad_cases.npz")
    np.savez_compressed(path, **{f"{k}__{f}": v[f] for k, v in out.items() for f in v})
    print(path)


if __name__ == "__main__":
    main()
