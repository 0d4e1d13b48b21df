"""The C-ABI library loads and exports every symbol include/tcb200.h declares (CPU, no compute),
and the product path fails loudly without a CUDA device."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

import tensorcircuit_b200 as tc
from tensorcircuit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "tcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tcb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared_functions()
    assert len(names) >= 19
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in tcb200.h is not exported" % n
    assert sorted(_lib.EXPORTS) == names


def test_version_and_errors():
    assert "sm_100a" in _lib.version()
    # argument validation happens before any CUDA call
    bits = np.array([0], dtype=np.int32)
    m = np.eye(2, dtype=np.complex128)
    rc = _lib.lib.tcb200_apply_dense(None, 3, 0, 1, _lib.iptr(bits), _lib.dptr(m.view(np.float64)), 1, None)
    assert rc == -1 and b"NULL" in _lib.lib.tcb200_last_error()
    with pytest.raises(_lib.EngineError):
        _lib.check(rc)
    buf = ctypes.create_string_buffer(64)
    bad = np.array([3], dtype=np.int32)
    rc = _lib.lib.tcb200_apply_dense(ctypes.cast(buf, ctypes.c_void_p), 3, 0, 1, _lib.iptr(bad), _lib.dptr(m.view(np.float64)), 1, None)
    assert rc == -1 and b"out of range" in _lib.lib.tcb200_last_error()
    rc = _lib.lib.tcb200_apply_dense(ctypes.cast(buf, ctypes.c_void_p), 3, 7, 1, _lib.iptr(bits), _lib.dptr(m.view(np.float64)), 1, None)
    assert rc == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    c = tc.Circuit(2)
    c.h(0)
    with pytest.raises(_lib.EngineError, match="no CPU execution path"):
        c.state()
    with pytest.raises(_lib.EngineError):
        c.expectation_ps(z=[0])
    with pytest.raises(_lib.EngineError):
        c.sample(batch=2, allow_state=True, status=[0.1, 0.2], format="sample_int")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensorcircuit_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), "%s imports the oracle" % f
                assert "sv_port" not in s and "tc_oracle" not in s and "libsvport" not in s, "%s uses the oracle" % f
