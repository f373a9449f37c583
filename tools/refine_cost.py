"""A/B of the FP64 refinement (kbo_set_tc_refine) at cfg3: sweep time with it on and off, alternating, same box."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kubeflow_b200.gp import GPEngine  # noqa: E402


def main():
    for shape in ((1024, 1 << 17, 8), (8192, 1 << 20, 32)):
        run(*shape)


def run(N, M, D):
    X = np.random.default_rng(1234).random((N, D))
    y = np.sin(3.0 * X.sum(axis=1) / np.sqrt(D)) + 0.1 * np.random.default_rng(1235).standard_normal(N)
    Xc = torch.tensor(np.random.default_rng(4321).random((M, D)), device="cuda")
    out = {"shape": [N, M, D]}
    engs = {r: GPEngine(0, kernel="matern52", length_scale=0.3 * np.sqrt(D), amplitude=1.0, noise=1e-3, var_mode="tc", tc_refine=r)
            for r in (True, False)}
    for e in engs.values():
        e.tell(X, y)
        e.ask(Xc)
    ts = {True: [], False: []}
    for _ in range(6):
        for r, e in engs.items():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); best = e.ask(Xc); b.record(); torch.cuda.synchronize()
            ts[r].append(a.elapsed_time(b))
            out[f"best_{r}"] = (best.index, best.value)
    out["sweep_ms_refine_on"] = ts[True]
    out["sweep_ms_refine_off"] = ts[False]
    out["contenders"] = engs[True].last_contenders()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
