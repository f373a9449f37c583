"""kbo_fit_append against a refit: time per added trial at a few history sizes (fixed θ, Matérn-5/2, D = 32)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kubeflow_b200.gp import GPEngine  # noqa: E402


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


def main():
    D, out = 32, []
    for N0, mode in ((500, "auto"), (2050, "tc"), (8130, "tc")):
        k = 8
        X = np.random.default_rng(1234).random((N0 + k, D))
        y = np.sin(3.0 * X.sum(axis=1) / np.sqrt(D)) + 0.1 * np.random.default_rng(1235).standard_normal(N0 + k)
        e = GPEngine(0, kernel="matern52", length_scale=0.3 * np.sqrt(D), amplitude=1.0, noise=1e-3, var_mode=mode)
        Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
        e.tell(Xd[:N0], yd[:N0]); e.tell(Xd[:N0], yd[:N0])
        refit = [timed(lambda: e.tell(Xd[:N0 + 1], yd[:N0 + 1])) for _ in range(3)]
        e.tell(Xd[:N0], yd[:N0])
        app = [timed(lambda i=i: e.append(X[i], y[i])) for i in range(N0, N0 + k)]
        out.append({"N": N0, "var_mode": mode, "refit_ms": float(np.median(refit)), "append_ms": float(np.median(app)), "append_first_ms": app[0]})
        e.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
