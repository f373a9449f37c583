"""Phase timings of kbo_fit (KBO_FIT_TRACE=1 prints gram / potrf / trtri / rest on stderr) at a few history sizes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
os.environ["KBO_FIT_TRACE"] = "1"
from kubeflow_b200.gp import GPEngine  # noqa: E402

D = 32
for N in (64, 512, 2048, 8192):
    X = np.random.default_rng(1234).random((N, D))
    y = np.sin(3.0 * X.sum(axis=1) / np.sqrt(D)) + 0.1 * np.random.default_rng(1235).standard_normal(N)
    e = GPEngine(0, kernel="matern52", length_scale=0.3 * np.sqrt(D), amplitude=1.0, noise=1e-3, var_mode="f64")
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    for _ in range(3):
        e.tell(Xd, yd)
    print(N, e.fit_info()["lml"], file=sys.stderr)
    e.close()
