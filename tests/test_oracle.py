"""CPU: pin oracle/gp_oracle.py against the golden vectors produced by the real scikit-learn GPR
(oracle/make_golden.py) and against self-authored known-answer checks."""
import os

import numpy as np
import pytest

from oracle import gp_oracle as O


def _suggest(g, **over):
    kw = dict(kind=g["kind"], length_scale=g["length_scale"], amplitude=g["amplitude"], noise=g["noise"],
              acq=g["acq_kind"], xi=g["xi"], kappa=g["kappa"])
    kw.update(over)
    return O.suggest(g["X"], g["y"], g["Xc"], **kw)


def test_oracle_matches_sklearn_golden(golden):
    r = _suggest(golden)
    np.testing.assert_allclose(r["fit"]["L"], golden["L"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(r["fit"]["alpha"], golden["alpha"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(r["mu"], golden["mu"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r["std"], golden["std"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(r["acq"], golden["acq"], rtol=0, atol=1e-9)
    assert r["index"] == golden["index"]
    assert abs(r["fit"]["lml"] - golden["lml"]) < 1e-7 * max(1.0, abs(golden["lml"]))


def test_n1_closed_form():
    # one trial: mu_n(x) = k(x,x1)·y1/(1+s²) with y normalised to 0 => mean is y1; var = 1 − k²/(1+s²)
    x1 = np.array([[0.3, 0.6]]); y1 = np.array([1.7]); Xc = np.random.default_rng(0).random((50, 2))
    fit = O.gp_fit(x1, y1, kind="rbf", length_scale=0.5, noise=1e-3)
    mu, std = O.gp_predict(fit, Xc)
    k = O.kernel_matrix(Xc, x1, 0.5, "rbf")[:, 0]
    np.testing.assert_allclose(mu, 1.7, atol=1e-12)
    np.testing.assert_allclose(std ** 2, 1 - k * k / (1 + 1e-3), atol=1e-12)


def test_candidate_equals_training_point_variance():
    X, y, Xc = O.synthetic(40, 10, 3)
    th = O.theta_of_record(3)
    fit = O.gp_fit(X, y, kind="matern52", length_scale=th["length_scale"], noise=th["noise"])
    mu, std = O.gp_predict(fit, X[:5])
    # posterior variance at a training point is below the noise level, and tiny vs. the prior (1)
    assert np.all(std ** 2 <= th["noise"] * fit["y_std"] ** 2 * 1.0001)


def test_permutation_and_shard_invariance():
    X, y, Xc = O.synthetic(64, 300, 4)
    th = O.theta_of_record(4)
    kw = dict(kind="matern52", length_scale=th["length_scale"], noise=th["noise"])
    a = O.suggest(X, y, Xc, **kw)
    p = np.random.default_rng(5).permutation(64)
    b = O.suggest(X[p], y[p], Xc, **kw)
    np.testing.assert_allclose(a["acq"], b["acq"], atol=1e-10)
    # sharding the grid: max over shards with lowest-global-index tie-break == global first argmax
    best = None
    for r in range(4):
        s = O.suggest(X, y, Xc[r * 75:(r + 1) * 75], **kw)
        cand = (s["value"], -(r * 75 + s["index"]))
        best = cand if best is None or cand > best else best
    assert -best[1] == a["index"]


def test_duplicate_candidate_lowest_index_wins():
    X, y, Xc = O.synthetic(32, 128, 3)
    th = O.theta_of_record(3)
    kw = dict(kind="rbf", length_scale=th["length_scale"], noise=th["noise"])
    i = O.suggest(X, y, Xc, **kw)["index"]
    Xd = np.concatenate([Xc, Xc[i:i + 1], Xc[i:i + 1]])
    assert O.suggest(X, y, Xd, **kw)["index"] == i
    j = 0 if i != 0 else 1
    Xe = Xc.copy(); Xe[j] = Xc[i]
    assert O.suggest(X, y, Xe, **kw)["index"] == min(i, j)


def test_properties():
    X, y, Xc = O.synthetic(100, 500, 5)
    th = O.theta_of_record(5)
    r = O.suggest(X, y, Xc, kind="matern52", length_scale=th["length_scale"], noise=th["noise"])
    K = O.kernel_matrix(X, X, th["length_scale"], "matern52")
    assert np.allclose(K, K.T) and np.linalg.eigvalsh(K).min() > -1e-10
    L = r["fit"]["L"]
    np.testing.assert_allclose(L @ L.T, K + th["noise"] * np.eye(100), atol=1e-12)
    assert (r["std"] >= 0).all() and (r["acq"] >= 0).all()
    # EI -> 0 when sigma -> 0 and mu > y_opt
    assert O.acquisition(np.array([2.0]), np.array([1e-12]), 1.0)[0] < 1e-300 + 1e-12
    assert O.acquisition(np.array([2.0]), np.array([0.0]), 1.0)[0] == 0.0


def test_lml_gradient_matches_sklearn_golden():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lmlgrad_cases.npz"), allow_pickle=False)
    names = sorted({k.split("__")[0] for k in z.files})
    assert len(names) == 4
    for n in names:
        ls = z[f"{n}__ls"]
        lml, g = O.lml_and_grad(z[f"{n}__X"], z[f"{n}__y"], kind=str(z[f"{n}__kind"]), length_scale=ls if len(ls) > 1 else float(ls[0]),
                                amplitude=float(z[f"{n}__amp"]), noise=float(z[f"{n}__noise"]))
        assert abs(lml - float(z[f"{n}__lml"])) < 1e-9
        np.testing.assert_allclose(g, z[f"{n}__grad"], rtol=0, atol=1e-9)
