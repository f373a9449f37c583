"""The synthetic workload of record, for everything that runs on the GPU side: bench.py's
product arm, tools/, the gRPC benchmarks.  ``oracle/gp_oracle.py`` carries its own statement of the same formulas for the CPU
side (tests/test_workload.py checks the two agree bit for bit); nothing here imports ``oracle``.

    X  = rng(1234).random((N, D))                      trials in the normalised cube (skopt's transformed space)
    y  = sin(3·Σ_d x_d / √D) + 0.1·rng(1235).normal(N)
    Xc = rng(4321).random((M_total, D))[offset : offset + M]      candidate grid; a rank owns a contiguous row block
    θ  : amplitude 1, ℓ_d = 0.3·√D, noise 1e-3, ξ = 0.01, κ = 1.96   (fixed: parity is defined at fixed θ)
"""
from __future__ import annotations

import numpy as np


def trials(N: int, D: int):
    X = np.random.default_rng(1234).random((N, D))
    y = np.sin(3.0 * X.sum(axis=1) / np.sqrt(D)) + 0.1 * np.random.default_rng(1235).standard_normal(N)
    return X, y


def candidates(M: int, D: int, *, offset: int = 0, dtype=np.float32):
    """Rows [offset, offset + M) of rng(4321).random((·, D)) — the generator is advanced past the skipped rows in blocks, so a
    rank's shard equals the same slice of the full grid without materialising it."""
    r = np.random.default_rng(4321)
    left = offset
    while left > 0:
        b = min(left, 1 << 20)
        r.random((b, D))
        left -= b
    return r.random((M, D)).astype(dtype)


def theta_of_record(D: int) -> dict:
    return dict(length_scale=0.3 * np.sqrt(D), amplitude=1.0, noise=1e-3, xi=0.01, kappa=1.96)


def describe(N: int, M: int, D: int, gpus: int = 1, kernel: str = "matern52", acq: str = "ei") -> str:
    """The ``config.workload`` string, identical for the product arm and the reference arm of bench.py."""
    return (f"cfg3: GP({kernel}) N={N} trials, D={D}, {acq.upper()} over M={M} candidates per GPU x {gpus} GPU(s) (grid {gpus * M}), fixed theta "
            f"(amp 1, ls 0.3*sqrt(D), noise 1e-3), one suggestion (fit + sweep + first-index argmax) per step")
