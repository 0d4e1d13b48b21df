"""ctypes binding of ``libtcb200.so`` (include/tcb200.h).

The engine has no CPU path: importing this module without the built library raises, and every
call fails loudly when no CUDA device is usable."""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtcb200.so")

C64, C128 = 0, 1
MAX_K = 5
MAX_DIAG_K = 12
MAX_PASS_OPS = 16
MAX_PASS_K = 4
MAX_TERMS = 8
MAX_GATE_PASS_OPS = 256
ERR_CAPACITY = -4
GATE_PASS_MAT_ELEMS = 1280


class EngineError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        # build in tree when a compiler is around (CPU container / first use); never fall back
        from . import build as _build

        try:
            _build.build()
        except Exception as e:  # pragma: no cover
            raise ImportError(
                "tensorcircuit_b200: %s is missing and could not be built (%s). "
                "Run `python tensorcircuit_b200/build.py`; there is no CPU fallback." % (LIB_PATH, e)
            )
    lib = ctypes.CDLL(LIB_PATH)
    sig = {
        "tcb200_version": (c_char_p, []),
        "tcb200_last_error": (c_char_p, []),
        "tcb200_launch_count": (c_int64, []),
        "tcb200_tma_pass_count": (c_int64, []),
        "tcb200_init_zero": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p]),
        "tcb200_load_c128": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
        "tcb200_set_zero": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p]),
        "tcb200_copy_rows": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
        "tcb200_apply_dense": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_double), c_int64, c_void_p]),
        "tcb200_apply_dense_batched": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), c_void_p, c_int64, c_void_p]),
        "tcb200_apply_diag": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_double), c_void_p, c_int64, c_void_p]),
        "tcb200_apply_pass": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_void_p, c_int64, c_int, POINTER(c_int), c_int64, c_void_p]),
        "tcb200_apply_pass_host": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_double), c_int, POINTER(c_int), c_int64, c_void_p]),
        "tcb200_apply_gate_pass": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_double), c_int, POINTER(c_int), c_int64, POINTER(c_double), c_void_p]),
        "tcb200_gate_pass_batched_workspace_bytes": (c_size_t, [c_int, c_int64]),
        "tcb200_apply_gate_pass_batched": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_double), POINTER(c_int), c_int, POINTER(c_int), c_int64, c_void_p, c_size_t, POINTER(c_double), c_void_p]),
        "tcb200_gate_pass_info": (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_double), c_int, POINTER(c_int), POINTER(c_double)]),
        "tcb200_apply_rpass_host": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_double), c_int, POINTER(c_int), c_int64, c_void_p]),
        "tcb200_pass_tile_bits": (c_int, [c_int]),
        "tcb200_norm2": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
        "tcb200_reduce_workspace_bytes": (c_size_t, [c_int, c_int64]),
        "tcb200_probability_state": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
        "tcb200_masked_norm2_workspace_bytes": (c_size_t, []),
        "tcb200_masked_norm2": (c_int, [c_void_p, c_int, c_int, c_uint64, c_uint64, c_void_p, c_void_p, c_size_t, c_void_p]),
        "tcb200_probability": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p]),
        "tcb200_expect_pauli": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_uint64), POINTER(c_uint64), POINTER(c_int), c_int, POINTER(c_int), c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
        "tcb200_expect_workspace_bytes": (c_size_t, [c_int, c_int64]),
        "tcb200_expect_z_max_terms": (c_int, [c_int]),
        "tcb200_expect_z_min_bits": (c_int, [c_int]),
        "tcb200_expect_z_workspace_bytes": (c_size_t, [c_int, c_int64]),
        "tcb200_expect_z": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_uint64), c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
        "tcb200_expect_tile_bits": (c_int, [c_int]),
        "tcb200_expect_single_flip_max_terms": (c_int, []),
        "tcb200_expect_single_flip_workspace_bytes": (c_size_t, [c_int, c_int64]),
        "tcb200_expect_single_flip": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_uint64), POINTER(c_int), c_int, POINTER(c_int), c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
        "tcb200_coo_expectation_workspace_bytes": (c_size_t, [c_int64, c_int64]),
        "tcb200_coo_expectation": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
        "tcb200_apply_pauli_sum_workspace_bytes": (c_size_t, [c_int]),
        "tcb200_apply_pauli_sum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_uint64), POINTER(c_uint64), POINTER(c_double), c_int64, c_void_p, c_size_t, c_void_p]),
        "tcb200_csr_matvec": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_void_p]),
        "tcb200_transition_local_max_ops": (c_int, []),
        "tcb200_transition_local_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
        "tcb200_transition_local": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_double), c_void_p, c_void_p, c_size_t, c_void_p]),
        "tcb200_sample": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_double, c_double, c_void_p, c_size_t, c_void_p]),
        "tcb200_sample_workspace_bytes": (c_size_t, [c_int]),
        "tcb200_run_circuit_host": (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_double), c_int64, POINTER(c_double), POINTER(c_int64), c_void_p, c_size_t, c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTS = [
    "tcb200_version", "tcb200_last_error", "tcb200_launch_count", "tcb200_tma_pass_count", "tcb200_init_zero", "tcb200_load_c128",
    "tcb200_set_zero", "tcb200_copy_rows",
    "tcb200_apply_dense", "tcb200_apply_dense_batched", "tcb200_apply_diag", "tcb200_apply_pass", "tcb200_apply_pass_host", "tcb200_apply_rpass_host", "tcb200_apply_gate_pass", "tcb200_gate_pass_info", "tcb200_gate_pass_batched_workspace_bytes", "tcb200_apply_gate_pass_batched",
    "tcb200_pass_tile_bits", "tcb200_norm2", "tcb200_reduce_workspace_bytes", "tcb200_masked_norm2_workspace_bytes", "tcb200_masked_norm2", "tcb200_probability_state", "tcb200_probability",
    "tcb200_expect_pauli", "tcb200_expect_workspace_bytes", "tcb200_expect_tile_bits",
    "tcb200_expect_single_flip_max_terms", "tcb200_expect_single_flip_workspace_bytes", "tcb200_expect_single_flip", "tcb200_expect_z_max_terms", "tcb200_expect_z_min_bits", "tcb200_expect_z_workspace_bytes", "tcb200_expect_z", "tcb200_sample",
    "tcb200_sample_workspace_bytes", "tcb200_run_circuit_host",
    "tcb200_coo_expectation_workspace_bytes", "tcb200_coo_expectation", "tcb200_apply_pauli_sum_workspace_bytes", "tcb200_apply_pauli_sum",
    "tcb200_transition_local_max_ops", "tcb200_transition_local_workspace_bytes", "tcb200_transition_local", "tcb200_csr_matvec",
]


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError("tcb200 error %d: %s" % (rc, lib.tcb200_last_error().decode()))


def version() -> str:
    return lib.tcb200_version().decode()


def launch_count() -> int:
    return int(lib.tcb200_launch_count())


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(POINTER(c_int))


def dptr(a: np.ndarray):
    assert a.flags.c_contiguous
    return a.ctypes.data_as(POINTER(c_double))


def u64ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return a.ctypes.data_as(POINTER(c_uint64))


def i64ptr(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(POINTER(c_int64))
