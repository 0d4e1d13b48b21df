"""ctypes wrapper of oracle/sv_port.c -- TEST / BASELINE INFRASTRUCTURE ONLY (see sv_port.c)."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsvport.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "sv_port.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.svp_num_threads.restype = ctypes.c_int
    return _lib


def _sfx(a):
    return "c64" if a.dtype == np.complex64 else "c128"


def apply(psi: np.ndarray, n: int, bits, u: np.ndarray) -> None:
    """In place: matrix index bit j <-> amplitude bit bits[j] (ascending)."""
    b = np.ascontiguousarray(bits, dtype=np.int32)
    m = np.ascontiguousarray(u, dtype=np.complex128)
    rc = getattr(lib(), "svp_apply_" + _sfx(psi))(psi.ctypes.data_as(ctypes.c_void_p), n, len(b), b.ctypes.data_as(ctypes.c_void_p), m.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0


def init_zero(psi: np.ndarray, n: int) -> None:
    getattr(lib(), "svp_init_zero_" + _sfx(psi))(psi.ctypes.data_as(ctypes.c_void_p), n)


def expect(psi: np.ndarray, n: int, flip: int, sign: int, ny: int) -> complex:
    out = np.zeros(2)
    getattr(lib(), "svp_expect_" + _sfx(psi))(psi.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_uint64(flip), ctypes.c_uint64(sign), ny, out.ctypes.data_as(ctypes.c_void_p))
    return complex(out[0], out[1])


def sample(psi: np.ndarray, n: int, u: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros(u.shape[0], dtype=np.int64)
    rc = getattr(lib(), "svp_sample_" + _sfx(psi))(psi.ctypes.data_as(ctypes.c_void_p), n, u.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(u.shape[0]), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def run_gatelist(n: int, ops, dtype=np.complex64) -> np.ndarray:
    """Gate by gate, as the reference does (one pass over the state per recorded gate)."""
    from . import tc_oracle as orc

    psi = np.empty(2**n, dtype=dtype)
    init_zero(psi, n)
    for name, q, p in ops:
        u = orc.gate_matrix(name, **p)
        k = len(q)
        # oracle matrix is big-endian in the caller's qubit order; reorder to ascending bits
        bits = [n - 1 - x for x in q]
        order = np.argsort(bits)
        t = u.reshape([2] * (2 * k))
        # axis a (out) <-> qubit q[a] <-> bit bits[a]; want index bit j <-> sorted bits[j]: big-endian axis order = descending bits
        perm = list(order[::-1])
        t = np.transpose(t, perm + [k + x for x in perm])
        apply(psi, n, sorted(bits), t.reshape(2**k, 2**k))
    return psi
