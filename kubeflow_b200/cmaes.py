"""CMA-ES sampler on the GPU (libkbo ``kbo_cma_*``) — host-side mirror of goptuna's ``cmaes`` sampler that Katib's
``cmaes`` algorithm uses: ``ask()`` a population, ``tell()`` its fitness."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


class CmaEs:
    def __init__(self, mean, sigma: float, popsize: int | None = None, seed: int = 0, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("kubeflow_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        mean = np.ascontiguousarray(np.asarray(mean, dtype=np.float64))
        self.D = int(mean.shape[0])
        self.popsize = int(popsize) if popsize else 4 + int(3 * np.log(self.D))
        self.device = int(device)
        self.lib = L.load()
        self._h, self._c = C.c_void_p(), C.c_void_p()
        rc = self.lib.kbo_create(C.byref(self._h), self.device)
        if rc != L.KBO_OK:
            raise L.KboError(rc, "kbo_create failed")
        L.check(self.lib, self._h, self.lib.kbo_cma_create(self._h, C.byref(self._c), self.D, self.popsize,
                                                           mean.ctypes.data_as(C.POINTER(C.c_double)), float(sigma), int(seed)))
        self._X = torch.empty(self.popsize, self.D, dtype=torch.float64, device=f"cuda:{self.device}")

    def close(self):
        if getattr(self, "_c", None):
            self.lib.kbo_cma_destroy(self._c)
            self._c = C.c_void_p()
        if getattr(self, "_h", None):
            self.lib.kbo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def ask(self, z: torch.Tensor | None = None) -> torch.Tensor:
        """Population (popsize × D float64 CUDA tensor).  ``z``: optional popsize×D standard normals (CUDA float64)."""
        zp = None
        if z is not None:
            if not (z.is_cuda and z.dtype == torch.float64 and tuple(z.shape) == (self.popsize, self.D)):
                raise ValueError("z must be a float64 CUDA tensor of shape (popsize, D)")
            z = z.contiguous()
            zp = z.data_ptr()
        L.check(self.lib, self._h, self.lib.kbo_cma_ask(self._h, self._c, self._X.data_ptr(), zp, self._stream()))
        return self._X

    def tell(self, fitness):
        f = fitness if isinstance(fitness, torch.Tensor) else torch.tensor(np.asarray(fitness, dtype=np.float64))
        f = f.to(device=f"cuda:{self.device}", dtype=torch.float64).contiguous()
        if f.numel() != self.popsize:
            raise ValueError(f"fitness must have {self.popsize} entries")
        L.check(self.lib, self._h, self.lib.kbo_cma_tell(self._h, self._c, f.data_ptr(), self._stream()))

    def state(self, with_Y: bool = False) -> dict:
        D, lam = self.D, self.popsize
        out = dict(mean=np.empty(D), sigma=np.empty(1), C=np.empty((D, D)), p_sigma=np.empty(D), pc=np.empty(D), B=np.empty((D, D)),
                   d=np.empty(D))
        Y = np.empty((lam, D)) if with_Y else None
        gen = C.c_int64()
        ptr = lambda a: a.ctypes.data if a is not None else None
        L.check(self.lib, self._h, self.lib.kbo_cma_state(self._h, self._c, ptr(out["mean"]), ptr(out["sigma"]), ptr(out["C"]),
                                                          ptr(out["p_sigma"]), ptr(out["pc"]), ptr(out["B"]), ptr(out["d"]), ptr(Y),
                                                          C.byref(gen), self._stream()))
        out["sigma"] = float(out["sigma"][0])
        out["generation"] = gen.value
        if with_Y:
            out["Y"] = Y
        return out

    def run_synthetic(self, fitness: str, generations: int):
        """BASELINE config 4: `generations` full generations on the device with a built-in fitness."""
        kind = {"sphere": 0, "rastrigin": 1}[fitness]
        best, ms, sw = C.c_double(), C.c_float(), C.c_double()
        L.check(self.lib, self._h, self.lib.kbo_cma_run_synthetic(self._h, self._c, kind, int(generations), C.byref(best), C.byref(ms),
                                                                  C.byref(sw)))
        return dict(best_f=best.value, elapsed_ms=ms.value, generations_per_s=generations / (ms.value * 1e-3), jacobi_sweeps=sw.value)
