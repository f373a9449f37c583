"""Small end-to-end invocations of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kubeflow_b200.gp import GPEngine
from kubeflow_b200.cmaes import CmaEs


def synthetic(N, M, D):
    """The workload of record, restated here so that tools/ never touches oracle/."""
    X = np.random.default_rng(1234).random((N, D))
    y = np.sin(3.0 * X.sum(axis=1) / np.sqrt(D)) + 0.1 * np.random.default_rng(1235).standard_normal(N)
    Xc = np.random.default_rng(4321).random((M, D))
    return X, y, Xc


def theta_of_record(D):
    return dict(length_scale=0.3 * np.sqrt(D), amplitude=1.0, noise=1e-3, xi=0.01, kappa=1.96)


for N, M, D, mode in ((70, 300, 3, "f64"), (135, 515, 5, "tc"), (300, 700, 33, "tc"), (700, 1900, 40, "tc"),   # 700: 2 tile pairs -> pruning pass
                      (1100, 2500, 12, "tc")):   # 1100: lazy inverse (leading 512 rows of W), alpha and survivors by the cooperative panel solves, 5 panels of the look-ahead factorisation
    X, y, Xc = synthetic(N, M, D)
    th = theta_of_record(D)
    e = GPEngine(0, kernel="matern52", acq="ei", var_mode=mode, **th)
    e.tell(X, y)
    b = e.ask(Xc.astype(np.float32))
    b2, t = e.suggest_host(X, y, Xc)
    e.tell(X[:N - 3], y[:N - 3])                 # bordered-Cholesky appends, a rebase, the FP64 refinement of the tc pick, LML gradient
    for i in range(N - 3, N):
        e.append(X[i], y[i])
    b3 = e.ask(Xc)
    e.rebase(N - 5, y[:N - 5] * 0.5)
    b4 = e.ask(Xc)
    g = e.lml_grad()[1]
    lm = e.lml_batch(X, y, [dict(noise=1e-3), dict(noise=1e-2, length_scale=th["length_scale"] * 2)])
    b5, _, _, _ = e.ask(Xc, return_arrays=True)      # three-product / FP64 array path
    print(N, M, D, mode, b.index, b2.index, b3.index, b4.index, b5.index, e.last_contenders(), e.last_prefix_survivors(), float(g[0]), lm.tolist())
    e.close()
    if mode == "tc":                                  # the ranking pass without pruning, and the round-1 ranking pass
        for kw in (dict(rank_prefix=0), dict(rank_tc=False)):
            e = GPEngine(0, kernel="matern52", acq="ei", var_mode=mode, **th, **kw)
            e.tell(X, y)
            print("   ", kw, e.ask(Xc.astype(np.float32)).index)
            e.close()
es = CmaEs(np.zeros(9), 1.0, popsize=20, seed=1)
print(es.run_synthetic("sphere", 3))
es.close()
