"""Time GPEngine.tell (fit only) at a few N; run once plain and once with KBO_FIT_SERIAL=1 for the A/B of the interleaved
Cholesky / row-panel inverse (fit.cu factor_and_invert).  KBO_FIT_TRACE=1 additionally prints the phase split."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kubeflow_b200 import workload as W
from kubeflow_b200.gp import GPEngine

out = {"serial": os.environ.get("KBO_FIT_SERIAL") is not None}
for N in (1024, 2048, 4096, 8192):
    D = 32
    X, y = W.trials(N, D)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    eng = GPEngine(0, kernel="matern52", acq="ei", var_mode="tc", **W.theta_of_record(D))
    for _ in range(3):
        eng.tell(Xd, yd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.tell(Xd, yd)
    e1.record()
    torch.cuda.synchronize()
    out[f"fit_ms_N{N}"] = e0.elapsed_time(e1) / 5
    out[f"lml_N{N}"] = eng.fit_info()["lml"]
    eng.close()
print(json.dumps(out))
