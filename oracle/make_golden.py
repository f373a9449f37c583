"""Generate tests/golden/gp_*.npz by running the REAL scikit-learn GaussianProcessRegressor
(1.9.0, importable in the authoring container; scikit-optimize itself is not) at fixed θ plus
scipy.stats.norm for the acquisition — i.e. the third-party code skopt's GP path executes
(SURVEY.md §8(c) item 1).  The vectors pin ``oracle/gp_oracle.py`` (tests/test_oracle.py) and are
the fixtures the GPU parity tests compare against.  Run:  python -m oracle.make_golden
"""
from __future__ import annotations

import os
import numpy as np
from scipy.stats import norm
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sk_suggest(X, y, Xc, kind, length_scale, amplitude, noise, acq, xi, kappa):
    base = RBF(length_scale=length_scale) if kind == "rbf" else Matern(length_scale=length_scale, nu=2.5)
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(amplitude, "fixed") * base, alpha=noise,
                                   normalize_y=True, optimizer=None)
    gpr.fit(X, y)
    mu, std = gpr.predict(Xc, return_std=True)
    y_opt = np.min(y)
    if acq == "lcb":
        a = -(mu - kappa * std)
    else:
        a = np.zeros_like(mu)
        mask = std > 0
        imp = y_opt - xi - mu[mask]
        z = imp / std[mask]
        a[mask] = imp * norm.cdf(z) + std[mask] * norm.pdf(z) if acq == "ei" else norm.cdf(z)
    return dict(mu=mu, std=std, acq=a, index=int(np.argmin(-a)), L=gpr.L_, alpha=gpr.alpha_,
                y_mean=gpr._y_train_mean, y_std=gpr._y_train_std,
                lml=gpr.log_marginal_likelihood(gpr.kernel_.theta))


def cases():
    r = np.random.default_rng(20260921)
    out = []

    def add(name, N, M, D, kind, acq, ls=None, amp=1.0, noise=1e-3, xi=0.01, kappa=1.96, mod=None):
        X = r.random((N, D))
        y = np.sin(3.0 * X.sum(1) / np.sqrt(D)) + 0.1 * r.standard_normal(N)
        Xc = r.random((M, D))
        ls = 0.3 * np.sqrt(D) if ls is None else ls
        if mod:
            X, y, Xc = mod(X, y, Xc)
        out.append(dict(name=name, X=X, y=y, Xc=Xc, kind=kind, acq=acq, length_scale=np.asarray(ls, float),
                        amplitude=amp, noise=noise, xi=xi, kappa=kappa))

    add("matern_ei_n48_d4", 48, 512, 4, "matern52", "ei")
    add("rbf_ei_n64_d3", 64, 384, 3, "rbf", "ei")
    add("matern_lcb_n33_d2", 33, 257, 2, "matern52", "lcb")
    add("rbf_pi_n40_d4", 40, 300, 4, "rbf", "pi")
    add("matern_ei_ard_n50_d3", 50, 200, 3, "matern52", "ei", ls=[0.4, 0.9, 1.7], amp=2.5, noise=1e-2)
    add("rbf_ei_n1_d2", 1, 64, 2, "rbf", "ei")                                   # N=1 closed form
    add("matern_ei_n2_d1", 2, 50, 1, "matern52", "ei")
    # duplicate candidates: rows 7 and 130 identical to row 3 -> lowest index must win if it is the max
    add("matern_ei_dupcand_n32_d3", 32, 160, 3, "matern52", "ei",
        mod=lambda X, y, Xc: (X, y, np.concatenate([Xc[:7], Xc[3:4], Xc[8:130], Xc[3:4], Xc[131:]])))
    # candidate == training point (σ² collapses to the noise-only level)
    add("rbf_ei_cand_is_train_n24_d2", 24, 100, 2, "rbf", "ei",
        mod=lambda X, y, Xc: (X, y, np.concatenate([X[:10], Xc[10:]])))
    # constant y: sklearn's zero-std guard (y_std -> 1)
    add("matern_ei_consty_n16_d2", 16, 90, 2, "matern52", "ei", mod=lambda X, y, Xc: (X, np.full_like(y, 0.75), Xc))
    # a mid-size case that crosses kernel tile boundaries (N, M not multiples of anything)
    add("matern_ei_n200_d8", 200, 1111, 8, "matern52", "ei")
    add("rbf_lcb_n130_d5", 130, 777, 5, "rbf", "lcb")
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for c in cases():
        res = sk_suggest(c["X"], c["y"], c["Xc"], c["kind"], c["length_scale"], c["amplitude"], c["noise"],
                         c["acq"], c["xi"], c["kappa"])
        path = os.path.join(OUT, f"gp_{c['name']}.npz")
        np.savez_compressed(path, X=c["X"], y=c["y"], Xc=c["Xc"], kind=c["kind"], acq_kind=c["acq"],
                            length_scale=c["length_scale"], amplitude=c["amplitude"], noise=c["noise"],
                            xi=c["xi"], kappa=c["kappa"], mu=res["mu"], std=res["std"], acq=res["acq"],
                            index=res["index"], alpha=res["alpha"], L=res["L"], y_mean=res["y_mean"],
                            y_std=res["y_std"], lml=res["lml"])
        print(f"{path}: N={len(c['y'])} M={len(c['Xc'])} argmax={res['index']} acq*={res['acq'][res['index']]:.6g}")


if __name__ == "__main__":
    main()
