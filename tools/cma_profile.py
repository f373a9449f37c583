import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from kubeflow_b200.cmaes import CmaEs
es = CmaEs(np.full(128, 3.0), 2.0, popsize=4096, seed=7)
print(es.run_synthetic("rastrigin", 12))
