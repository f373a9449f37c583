"""Helper of test_fit_variants_agree: one fit + one suggestion in a fresh process (the KBO_FIT_* switches are read once per process),
L, W's leading block, alpha and the suggestion written to an .npz."""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import gp_oracle as O            # noqa: E402  (tests/ may use the oracle's workload generator)
from kubeflow_b200.gp import GPEngine        # noqa: E402

N, M, D, mode, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
X, y, Xc = O.synthetic(N, M, D)
th = O.theta_of_record(D)
e = GPEngine(0, kernel="matern52", acq="ei", var_mode=mode, length_scale=th["length_scale"], amplitude=th["amplitude"], noise=th["noise"],
             xi=th["xi"], kappa=th["kappa"])
e.tell(X, y)
b = e.ask(Xc)
info = e.fit_info()
Lm, Wm, al = e.state()
np.savez(out, L=Lm.cpu().numpy(), W=Wm.cpu().numpy(), alpha=al.cpu().numpy(), lml=info["lml"], index=b.index, value=b.value, mu=b.mu, std=b.std)
e.close()
