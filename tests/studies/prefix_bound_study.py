"""Study (CPU, oracle arithmetic): how well does the PREFIX bound prune?  sigma²_prefix(J) = amp − Σ_{j < J} v_j² (the variance
explained by the first J trials only) is an upper bound on sigma², so EI(mu, sigma²_prefix) >= EI(mu, sigma²).  For the workload
of record: how many candidates keep an upper bound above (a) the best exact EI of a stratified sample (what the calibration rows
give for free) and (b) the true maximum, for prefixes of 1/8, 1/4, 1/2 of the trials (cost of the triangular contraction up to
that prefix: 1/64, 1/16, 1/4 of the whole)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scipy.linalg import solve_triangular
from oracle import gp_oracle as O

N, M, D = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 40000, 32)))
X, y, Xc = O.synthetic(N, M, D)
th = O.theta_of_record(D)
fit = O.gp_fit(X, y, kind="matern52", **{k: th[k] for k in ("length_scale", "amplitude", "noise")})
Ks = O.kernel_matrix(Xc, X, th["length_scale"], "matern52", th["amplitude"])
mu = fit["y_std"] * (Ks @ fit["alpha"]) + fit["y_mean"]
V = solve_triangular(fit["L"], Ks.T, lower=True, check_finite=False)      # N × M
cum = np.cumsum(V * V, axis=0)
yopt = float(y.min())
full = O.acquisition(mu, np.sqrt(np.maximum(th["amplitude"] - cum[-1], 0) * fit["y_std"] ** 2), yopt, "ei", th["xi"], th["kappa"])
best = full.max()
sample = full[:: max(1, M // 400)]       # a stratified sample as sparse (relative to M) as 18944 of 1M
print(f"N={N} M={M} D={D}: max EI {best:.4f}, best of a {len(sample)}-row stratified sample {sample.max():.4f}, mean {full.mean():.2e}")
for frac in (0.125, 0.25, 0.5):
    J = int(N * frac)
    var_ub = np.maximum(th["amplitude"] - cum[J - 1], 0)
    ub = O.acquisition(mu, np.sqrt(var_ub * fit["y_std"] ** 2), yopt, "ei", th["xi"], th["kappa"])
    assert (ub >= full - 1e-12).all()
    print(f"  prefix {frac:5.3f} (cost {frac ** 2:.3f}): survivors vs sample-best {int((ub >= sample.max()).sum()):6d}  vs true max {int((ub >= best).sum()):6d}"
          f"  median sigma²_prefix/sigma² {np.median(var_ub / np.maximum(th['amplitude'] - cum[-1], 1e-12)):.3f}")
