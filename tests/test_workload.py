"""CPU: the product-side statement of the workload of record (kubeflow_b200.workload) and the oracle's agree bit for bit."""
import numpy as np

from kubeflow_b200 import workload as W
from oracle import gp_oracle as O


def test_workload_matches_the_oracle_statement():
    X, y, Xc = O.synthetic(300, 5000, 7)
    Xw, yw = W.trials(300, 7)
    np.testing.assert_array_equal(X, Xw)
    np.testing.assert_array_equal(y, yw)
    np.testing.assert_array_equal(Xc, W.candidates(5000, 7, dtype=np.float64))
    _, _, part = O.synthetic(300, 1200, 7, m_offset=2100, m_total=5000)
    np.testing.assert_array_equal(part, W.candidates(1200, 7, offset=2100, dtype=np.float64))
    assert W.theta_of_record(7) == O.theta_of_record(7)
    assert W.describe(8192, 1048576, 32, 2) == W.describe(8192, 1048576, 32, 2) and "x 2 GPU" in W.describe(8192, 1048576, 32, 2)
