"""CPU: known-answer / property checks pinning oracle/cma_oracle.py (no upstream golden vectors exist offline)."""
import numpy as np

from oracle import cma_oracle as C


def test_parameters_match_tutorial_defaults():
    p = C.CmaParams(10, 4 + int(3 * np.log(10)))          # λ = 10 for n = 10 (tutorial eq. 48)
    assert p.popsize == 10 and p.mu == 5
    assert abs(p.weights[:p.mu].sum() - 1.0) < 1e-12 and (p.weights[:p.mu] > 0).all() and (p.weights[p.mu:] <= 0).all()
    assert np.all(np.diff(p.weights[:p.mu]) < 0)
    assert abs(p.mu_eff - 1.0 / (p.weights[:p.mu] ** 2).sum()) < 1e-12     # μ_eff = 1/Σw² for normalised positive weights
    assert 0 < p.c1 < p.cmu < 1 and p.c1 + p.cmu <= 1
    assert abs(p.chi_n - np.sqrt(10) * (1 - 1 / 40 + 1 / 2100)) < 1e-15
    # negative weights sum to -min(alpha_mu, alpha_mueff, alpha_posdef): keeps C positive definite (tutorial eq. 53)
    assert p.c1 + p.cmu * p.weights.sum() >= -1e-12


def test_update_is_rotation_equivariant():
    """CMA-ES is invariant under orthogonal transforms of the search space: rotating the steps rotates the state."""
    r = np.random.default_rng(0)
    n, lam = 6, 12
    Q, _ = np.linalg.qr(r.standard_normal((n, n)))
    z = r.standard_normal((lam, n))
    f = r.standard_normal(lam)
    a, b = C.CmaState(np.zeros(n), 0.7, lam), C.CmaState(np.zeros(n), 0.7, lam)
    _, Ya = C.ask(a, z)
    C.tell(a, Ya, f)
    C.tell(b, Ya @ Q.T, f)
    np.testing.assert_allclose(b.mean, Q @ a.mean, atol=1e-12)
    np.testing.assert_allclose(b.C, Q @ a.C @ Q.T, atol=1e-12)
    np.testing.assert_allclose(b.p_sigma, Q @ a.p_sigma, atol=1e-12)
    assert abs(a.sigma - b.sigma) < 1e-12


def test_converges_on_sphere_and_keeps_C_spd():
    r = np.random.default_rng(1)
    n, lam = 8, 16
    s = C.CmaState(np.full(n, 3.0), 2.0, lam)
    for g in range(150):
        X, Y = C.ask(s, r.standard_normal((lam, n)))
        C.tell(s, Y, C.sphere(X))
        assert np.linalg.eigvalsh((s.C + s.C.T) / 2).min() > 0
    assert C.sphere(s.mean[None])[0] < 1e-8 and s.sigma < 1e-3


def test_ties_broken_by_sample_index():
    s1, s2 = C.CmaState(np.zeros(3), 1.0, 6), C.CmaState(np.zeros(3), 1.0, 6)
    z = np.random.default_rng(2).standard_normal((6, 3))
    _, Y = C.ask(s1, z)
    C.tell(s1, Y, np.array([1.0, 0.0, 0.0, 2.0, 0.0, 3.0]))
    C.tell(s2, Y[[1, 2, 4, 0, 3, 5]], np.array([0.0, 0.0, 0.0, 1.0, 2.0, 3.0]))
    np.testing.assert_allclose(s1.mean, s2.mean, atol=1e-15)
