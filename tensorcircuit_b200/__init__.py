"""tensorcircuit_b200 -- a B200-native statevector engine behind TensorCircuit's circuit API.

``import tensorcircuit_b200 as tc`` gives the hot-path surface of the reference
(tensorcircuit/__init__.py:1-72): ``tc.Circuit``, ``tc.gates``, ``tc.backend``,
``tc.set_backend / set_dtype / set_contractor``, ``tc.quantum``, ``tc.templates``.
Everything O(2^n) runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/tcb200.h``; there is no CPU, jax, tensorflow or torch-op execution path."""

__version__ = "0.1.0"
__author__ = "tensorcircuit_b200"

from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .cons import (  # noqa: F401
    contractor,
    dtypestr,
    get_backend,
    get_contractor,
    get_dtype,
    npdtype,
    rdtypestr,
    runtime_backend,
    runtime_contractor,
    runtime_dtype,
    set_backend,
    set_contractor,
    set_distributed,
    set_dtype,
    set_function_backend,
    set_function_contractor,
    set_function_dtype,
)
from . import b200_backend as _backend_module  # noqa: F401  binds cons.backend
from . import gates  # noqa: F401
from .gates import Gate, array_to_tensor, num_to_tensor  # noqa: F401
from . import quantum  # noqa: F401
from . import channels  # noqa: F401
from .circuit import Circuit, DeviceArray, expectation  # noqa: F401
from . import templates  # noqa: F401
from . import engine  # noqa: F401
from . import parallel  # noqa: F401

backend = _backend_module.get_backend()

# Recording a circuit allocates a few small objects per gate; with torch and numpy loaded, a full (generation-2)
# cyclic collection walks their whole module heap and stalls the host for 50-170 ms every few thousand gate calls --
# longer than all kernels of a 30-qubit step together.  Everything imported so far is module-level and lives for the
# whole process anyway: move it to the permanent generation (TCB200_GC_FREEZE=0 to leave the collector alone).
import gc as _gc  # noqa: E402
import os as _os  # noqa: E402

if _os.environ.get("TCB200_GC_FREEZE", "1") != "0":
    _gc.collect()
    _gc.freeze()

# A vmap batch of 1024 parameter sets records one [1024, 4, 4] complex128 matrix per two-qubit gate (256 KiB) and keeps
# it until the flush.  glibc serves every request above 128 KiB with a fresh mmap, so each of those arrays is 64 first-touch
# page faults (2-3 us each on a virtualised host) and a munmap afterwards: 110 of the 200 us a batched rzz costs, half of
# the host time of config 3.  Keep requests up to 32 MiB on the heap and do not trim it, so the pages of one call are
# reused by the next (TCB200_MALLOPT=0 to leave the allocator alone).
if _os.environ.get("TCB200_MALLOPT", "1") != "0":
    try:
        import ctypes as _ctypes

        _libc = _ctypes.CDLL("libc.so.6")
        _libc.mallopt(-3, 32 << 20)  # M_MMAP_THRESHOLD (32 MiB is the largest value glibc accepts)
        _libc.mallopt(-1, 1 << 30)  # M_TRIM_THRESHOLD
        _libc.mallopt(-2, 64 << 20)  # M_TOP_PAD
    except (OSError, AttributeError):  # not glibc: nothing to tune
        pass


def about() -> None:
    """tensorcircuit/about.py: versions of what the engine runs on"""
    import platform

    import numpy
    import torch

    print("tensorcircuit_b200 %s  (%s)" % (__version__, _lib.version()))
    print("python %s, numpy %s, torch %s, cuda available: %s" % (platform.python_version(), numpy.__version__, torch.__version__, torch.cuda.is_available()))


def __getattr__(name):  # live view of the mutable globals (set_dtype rebinds them in cons)
    from . import cons as _cons

    if name in ("dtypestr", "rdtypestr", "npdtype"):
        return getattr(_cons, name)
    raise AttributeError(name)
