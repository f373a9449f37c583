"""CPU: the C-ABI library builds/loads and exports every symbol include/kbo.h declares (no compute calls)."""
import ctypes
import os
import re

from kubeflow_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "kbo.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kbo_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"libkbo.so does not export {s}"
    assert sorted(_lib.EXPORTS) == syms, "kubeflow_b200/_lib.py EXPORTS out of sync with include/kbo.h"
    assert lib.kbo_version() == 100


def test_struct_layouts_match_header():
    # sizes the C side uses (kbo_params: 4*i32 + 4*f64 + ptr + 2*i32 = 64; kbo_best 32; kbo_timings 44)
    assert ctypes.sizeof(_lib.KboParams) == 64
    assert ctypes.sizeof(_lib.KboBest) == 32
    assert ctypes.sizeof(_lib.KboTimings) == 44


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.kbo_create(ctypes.byref(h), 0) == _lib.KBO_ERR_CUDA
    assert not h
    import pytest
    from kubeflow_b200.gp import GPEngine
    with pytest.raises(RuntimeError):
        GPEngine(0)
