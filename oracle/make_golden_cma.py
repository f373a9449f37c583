"""Generates the CMA-ES fixtures under tests/golden/ (run once; the outputs are committed):

  cma_known_answers.json   strategy parameters for (n, λ) = (10, 10), (2, 6), (128, 4096) evaluated from the formulas of
                           N. Hansen, "The CMA Evolution Strategy: A Tutorial" (arXiv:1604.00772) — eqs. (49)–(58), Table 1 — with
                           Python's `decimal` at 50 digits, i.e. independently of NumPy and of oracle/cma_oracle.py's code path.
  cma_replay.npz           a seeded 12-generation run of oracle/cma_oracle.py on the sphere and on Rastrigin (n = 5, λ = 8):
                           mean, σ, C, paths after every generation, so that any drift of the oracle itself is caught bit for bit.

PARITY UNPINNED BY THE REFERENCE: neither file comes from goptuna / `cmaes` (not installable here); they pin the oracle to
the published formulas and to itself."""
import json
import os
import sys
from decimal import Decimal as Dc, getcontext

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cma_oracle as C  # noqa: E402

getcontext().prec = 50
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def params(n, lam):
    mu = lam // 2
    wp = [(Dc(lam + 1) / 2).ln() - Dc(i).ln() for i in range(1, lam + 1)]
    s1 = sum(wp[:mu])
    mu_eff = s1 * s1 / sum(w * w for w in wp[:mu])
    s2 = sum(wp[mu:])
    mu_eff_minus = s2 * s2 / sum(w * w for w in wp[mu:])
    n_ = Dc(n)
    c1 = Dc(2) / ((n_ + Dc("1.3")) ** 2 + mu_eff)
    cmu = min(1 - c1 - Dc("1e-8"), Dc(2) * (mu_eff - 2 + 1 / mu_eff) / ((n_ + 2) ** 2 + Dc(2) * mu_eff / 2))
    a_mu = 1 + c1 / cmu
    a_mueff = 1 + 2 * mu_eff_minus / (mu_eff + 2)
    a_pd = (1 - c1 - cmu) / (n_ * cmu)
    min_alpha = min(a_mu, a_mueff, a_pd)
    pos = sum(w for w in wp if w > 0)
    neg = sum(-w for w in wp if w < 0)
    weights = [w / pos if w >= 0 else min_alpha / neg * w for w in wp]
    c_sigma = (mu_eff + 2) / (n_ + mu_eff + 5)
    root = ((mu_eff - 1) / (n_ + 1)).sqrt() - 1
    d_sigma = 1 + 2 * max(Dc(0), root) + c_sigma
    cc = (4 + mu_eff / n_) / (n_ + 4 + 2 * mu_eff / n_)
    chi_n = n_.sqrt() * (1 - 1 / (4 * n_) + 1 / (21 * n_ * n_))
    return {"n": n, "popsize": lam, "mu": mu, "mu_eff": float(mu_eff), "c1": float(c1), "cmu": float(cmu), "c_sigma": float(c_sigma),
            "d_sigma": float(d_sigma), "cc": float(cc), "chi_n": float(chi_n), "weights_first8": [float(w) for w in weights[:8]],
            "weights_last2": [float(w) for w in weights[-2:]], "weights_sum": float(sum(weights))}


def replay():
    out = {}
    for name, f in (("sphere", C.sphere), ("rastrigin", C.rastrigin)):
        r = np.random.default_rng(20240917)
        st = C.CmaState(np.full(5, 1.5), 0.8, 8)
        means, sigmas, Cs, ps, pcs = [], [], [], [], []
        for _ in range(12):
            X, Y = C.ask(st, r.standard_normal((8, 5)))
            C.tell(st, Y, f(X))
            means.append(st.mean.copy()); sigmas.append(st.sigma); Cs.append(st.C.copy()); ps.append(st.p_sigma.copy()); pcs.append(st.pc.copy())
        out.update({f"{name}_mean": np.array(means), f"{name}_sigma": np.array(sigmas), f"{name}_C": np.array(Cs), f"{name}_p_sigma": np.array(ps),
                    f"{name}_pc": np.array(pcs)})
    return out


if __name__ == "__main__":
    json.dump({"source": "arXiv:1604.00772 eqs. (49)-(58), evaluated with decimal (50 digits) by oracle/make_golden_cma.py",
               "cases": [params(10, 10), params(2, 6), params(128, 4096)]}, open(os.path.join(OUT, "cma_known_answers.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "cma_replay.npz"), **replay())
    print("written")
