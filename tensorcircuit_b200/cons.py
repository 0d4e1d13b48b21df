"""Process-global configuration, mirroring tensorcircuit/cons.py for the hot path.

``set_dtype`` (cons.py:129-182) selects complex64 / complex128 for every Circuit created
afterwards.  ``set_backend`` (cons.py:36-86) always binds the one B200 backend: the engine has
no jax / tensorflow / torch-op / numpy execution path, so the reference's backend names are
accepted as aliases to let existing scripts run unchanged.  ``set_contractor``
(cons.py:732-830) is accepted and ignored -- there is no tensor network to contract."""

from __future__ import annotations

import logging
import sys
from contextlib import contextmanager
from typing import Any, Iterator, Optional

import numpy as np

logger = logging.getLogger(__name__)

package_name = "tensorcircuit_b200"
thismodule = sys.modules[__name__]

dtypestr = "complex64"
rdtypestr = "float32"
npdtype = np.complex64
backend: Any = None  # bound by b200_backend.py at import
contractor: Any = None
# True: under torchrun (world size > 1) every Circuit shards its state over the ranks
distributed_state = False

_BACKEND_ALIASES = ("b200", "cuda", "numpy", "jax", "tensorflow", "pytorch", "cupy")


def _rebind(name: str, value: Any) -> None:
    for mod in list(sys.modules):
        if mod.startswith(package_name):
            m = sys.modules[mod]
            if hasattr(m, name):
                setattr(m, name, value)


def set_backend(backend_name: Optional[str] = None, set_global: bool = True) -> Any:
    from .b200_backend import get_backend

    if backend_name is None:
        backend_name = "b200"
    if hasattr(backend_name, "name"):
        backend_name = backend_name.name
    if backend_name not in _BACKEND_ALIASES:
        raise ValueError("Backend '%s' does not exist" % backend_name)
    if backend_name not in ("b200", "cuda"):
        logger.info("backend '%s' requested: the B200 statevector backend is used instead", backend_name)
    b = get_backend()
    if set_global:
        _rebind("backend", b)
    return b


set_backend.__doc__ = "Bind the B200 backend (reference names are aliases)."


def set_dtype(dtype: Optional[str] = None, set_global: bool = True) -> Any:
    if not dtype:
        dtype = "complex64"
    if dtype == "complex64":
        rdtype = "float32"
    elif dtype == "complex128":
        rdtype = "float64"
    else:
        raise ValueError(f"Unsupported data type: {dtype}")  # cons.py:153
    npd = getattr(np, dtype)
    if set_global:
        _rebind("dtypestr", dtype)
        _rebind("rdtypestr", rdtype)
        _rebind("npdtype", npd)
    return dtype, rdtype


def get_dtype(dtype: Optional[str] = None) -> Any:
    """cons.py:180: the (complex, real) dtype names without touching the global setting"""
    return set_dtype(dtype, set_global=False)


def get_backend(backend_name: Optional[str] = None) -> Any:
    """backends/backend_factory.py:38-59: every name resolves to the one engine backend"""
    return set_backend(backend_name, set_global=False)


def get_contractor(method: Optional[str] = None, *args: Any, **kws: Any) -> Any:
    return None


def set_distributed(flag: bool = True, sample_order: Optional[str] = None) -> bool:
    """Shard every Circuit's state vector over the ranks of the default process group (top
    log2(G) index bits = rank; see tensorcircuit_b200.dist).  SPMD: all ranks run the same
    script and get identical host-side results.  ``sample_order="logical"`` makes ``c.sample(status=...)``
    return the single-GPU indices for the same uniforms (the state is brought back to the identity layout
    first: up to two more remaps); the default "physical" samples in place."""
    if sample_order is not None:
        if sample_order not in ("physical", "logical"):
            raise ValueError("sample_order must be 'physical' or 'logical'")
        from .dist import DistState

        DistState.sample_order = sample_order
    _rebind("distributed_state", bool(flag))
    return bool(flag)


def set_contractor(method: Optional[str] = None, *args: Any, **kws: Any) -> Any:
    """Accepted for source compatibility; the statevector engine has no contraction path."""
    return None


set_function_contractor = lambda *a, **k: (lambda f: f)  # noqa: E731


@contextmanager
def runtime_backend(backend_name: Optional[str] = None) -> Iterator[Any]:
    yield set_backend(backend_name, set_global=False)


@contextmanager
def runtime_dtype(dtype: Optional[str] = None) -> Iterator[Any]:
    old = thismodule.dtypestr
    r = set_dtype(dtype)
    yield r
    set_dtype(old)


@contextmanager
def runtime_contractor(*a: Any, **k: Any) -> Iterator[Any]:
    yield None


def set_function_backend(backend_name: Optional[str] = None):
    return lambda f: f


def set_function_dtype(dtype: Optional[str] = None):
    def wrapper(f):
        def newf(*args, **kws):
            with runtime_dtype(dtype):
                return f(*args, **kws)

        return newf

    return wrapper
