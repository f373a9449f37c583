"""One warm-up + N timed suggestions at cfg3 through kbo_suggest_host — the command profiled under ncu
(see profiles/README.md).  Prints the library's own CUDA-event phase timings for the last step."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kubeflow_b200.gp import GPEngine
from oracle import gp_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=8192)
ap.add_argument("--candidates", type=int, default=1_048_576)
ap.add_argument("--dim", type=int, default=32)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--var-mode", default="tc")
ap.add_argument("--k-span", type=int, default=0)
a = ap.parse_args()
X, y, _ = O.synthetic(a.trials, 1, a.dim)
Xc = np.random.default_rng(4321).random((a.candidates, a.dim)).astype(np.float32)
th = O.theta_of_record(a.dim)
eng = GPEngine(0, kernel="matern52", acq="ei", var_mode=a.var_mode, tc_k_span=a.k_span, **th)
for i in range(a.warmup + a.steps):
    best, t = eng.suggest_host(X, y, Xc)
print(json.dumps({"best": best.__dict__, "timings": t}))
