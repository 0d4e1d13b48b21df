"""Deferred expectation values.

The reference idiom for an energy is a Python loop over ``c.expectation_ps`` (one term per call:
examples/vqe_parallel_pmap.py:28-34, benchmarks/scripts/vqe_tc.py:109-141); under ``jit`` the
reference fuses those calls into one compiled program.  Here every call would otherwise be one
kernel launch + one device-to-host read.  Instead ``expectation_ps`` on an unbatched circuit
returns a :class:`LazyScalar`: the term is only *registered* with the circuit, linear arithmetic
on the result (``+ - *`` with numbers, ``K.real``) stays symbolic, and the first time a number
is really needed (``float()``, ``np.asarray``, a comparison, a print, a new gate on the
circuit ...) ALL terms registered so far on that state are evaluated together -- one launch
group, one read of the state per tile geometry, one device-to-host copy.  Results are
identical to the eager evaluation (same kernels, same float64 accumulation)."""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np


class TermPool:
    """Pending Pauli-string terms of one circuit state."""

    def __init__(self, circuit: Any):
        self.circuit = circuit
        self.fl: List[int] = []
        self.sg: List[int] = []
        self.ny: List[int] = []
        self.values: List[Optional[complex]] = []
        self.closed = False  # the circuit moved on: no new terms may join

    def add(self, fl: int, sg: int, ny: int) -> int:
        self.fl.append(fl)
        self.sg.append(sg)
        self.ny.append(ny)
        self.values.append(None)
        return len(self.values) - 1

    def flush(self) -> None:
        todo = [i for i, v in enumerate(self.values) if v is None]
        if not todo:
            return
        st = self.circuit._ensure_state()
        r = st.expectation_terms([self.fl[i] for i in todo], [self.sg[i] for i in todo], [self.ny[i] for i in todo])
        for k, i in enumerate(todo):
            self.values[i] = complex(r[0, k])

    def value(self, i: int) -> complex:
        if self.values[i] is None:
            self.flush()
        return self.values[i]  # type: ignore[return-value]


def _is_number(x: Any) -> bool:
    return isinstance(x, (int, float, complex, np.number)) or (isinstance(x, np.ndarray) and x.ndim == 0)


class LazyScalar:
    """``const + sum_i coef_i * <term_i>`` (mode 'c'), or its real / imaginary part (modes 're',
    'im').  Anything that is not linear forces the value."""

    __array_priority__ = 900
    shape: Tuple[int, ...] = ()
    ndim = 0
    size = 1

    def __init__(self, terms: List[Tuple[TermPool, int, complex]], const: complex = 0j, mode: str = "c", dtype: str = "complex64"):
        self._terms = terms
        self._const = const
        self._mode = mode
        self._dtype = dtype
        self._cached: Optional[Any] = None

    # -- evaluation --------------------------------------------------------------------------
    def _complex_value(self) -> complex:
        v = self._const
        for pool, i, c in self._terms:
            v = v + c * pool.value(i)
        return v

    def force(self) -> Any:
        if self._cached is None:
            v = self._complex_value()
            if self._mode == "c":
                self._cached = np.array(v, dtype=self._dtype)
            else:
                rd = np.float32 if self._dtype == "complex64" else np.float64
                self._cached = np.array(v.real if self._mode == "re" else v.imag, dtype=rd)
        return self._cached

    @property
    def dtype(self) -> Any:
        if self._mode == "c":
            return np.dtype(self._dtype)
        return np.dtype(np.float32 if self._dtype == "complex64" else np.float64)

    def __array__(self, dtype: Any = None, copy: Any = None) -> np.ndarray:
        a = self.force()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self) -> float:
        return float(np.real(self.force()))

    def __complex__(self) -> complex:
        return complex(self.force())

    def __int__(self) -> int:
        return int(float(self))

    def __bool__(self) -> bool:
        return bool(self.force())

    def __abs__(self) -> Any:
        return np.abs(self.force())

    def __repr__(self) -> str:
        return repr(self.force())

    def __format__(self, spec: str) -> str:
        return format(self.force()[()], spec)

    def item(self) -> Any:
        return self.force().item()

    def astype(self, dtype: Any) -> Any:
        return self.force().astype(dtype)

    def numpy(self) -> np.ndarray:
        return self.force()

    def reshape(self, *shape: Any) -> np.ndarray:
        return self.force().reshape(*shape)

    def __getitem__(self, idx: Any) -> Any:
        return self.force()[idx]

    def __hash__(self) -> int:
        return id(self)

    # -- linear structure ----------------------------------------------------------------------
    def _scaled(self, s: Any) -> Optional["LazyScalar"]:
        s = complex(s)
        if self._mode != "c" and s.imag != 0:
            return None
        return LazyScalar([(p, i, c * s) for p, i, c in self._terms], self._const * s, self._mode, self._dtype)

    def _added(self, o: Any, sign: float) -> Optional["LazyScalar"]:
        if isinstance(o, LazyScalar):
            if o._mode != self._mode:
                return None
            dt = "complex128" if "complex128" in (self._dtype, o._dtype) else "complex64"
            return LazyScalar(self._terms + [(p, i, sign * c) for p, i, c in o._terms], self._const + sign * o._const, self._mode, dt)
        if _is_number(o):
            oc = complex(o)
            if self._mode != "c":
                if oc.imag != 0:
                    return None
                # Re(z) + r = Re(z + r);  Im(z) + r = Im(z + i r)
                oc = oc if self._mode == "re" else 1j * oc.real
            return LazyScalar(list(self._terms), self._const + sign * oc, self._mode, self._dtype)
        return None

    def __add__(self, o: Any) -> Any:
        r = self._added(o, 1.0)
        return r if r is not None else self.force() + (o.force() if isinstance(o, LazyScalar) else o)

    __radd__ = __add__

    def __sub__(self, o: Any) -> Any:
        r = self._added(o, -1.0)
        return r if r is not None else self.force() - (o.force() if isinstance(o, LazyScalar) else o)

    def __rsub__(self, o: Any) -> Any:
        return (-self).__add__(o)

    def __neg__(self) -> "LazyScalar":
        return self._scaled(-1.0)  # type: ignore[return-value]

    def __pos__(self) -> "LazyScalar":
        return self

    def __mul__(self, o: Any) -> Any:
        if _is_number(o):
            r = self._scaled(o)
            if r is not None:
                return r
        return self.force() * (o.force() if isinstance(o, LazyScalar) else o)

    __rmul__ = __mul__

    def __truediv__(self, o: Any) -> Any:
        if _is_number(o) and complex(o) != 0:
            r = self._scaled(1.0 / complex(o))
            if r is not None:
                return r
        return self.force() / (o.force() if isinstance(o, LazyScalar) else o)

    def __rtruediv__(self, o: Any) -> Any:
        return o / self.force()

    def __pow__(self, p: Any) -> Any:
        return self.force() ** p

    def __rpow__(self, b: Any) -> Any:
        return b ** self.force()

    @property
    def real(self) -> Any:
        if self._mode == "c":
            return LazyScalar(list(self._terms), self._const, "re", self._dtype)
        return self

    @property
    def imag(self) -> Any:
        if self._mode == "c":
            return LazyScalar(list(self._terms), self._const, "im", self._dtype)
        return np.zeros((), dtype=self.dtype)

    def conj(self) -> Any:
        return self if self._mode != "c" else np.conj(self.force())

    conjugate = conj

    # comparisons force
    def __lt__(self, o: Any) -> Any:
        return self.force() < (o.force() if isinstance(o, LazyScalar) else o)

    def __le__(self, o: Any) -> Any:
        return self.force() <= (o.force() if isinstance(o, LazyScalar) else o)

    def __gt__(self, o: Any) -> Any:
        return self.force() > (o.force() if isinstance(o, LazyScalar) else o)

    def __ge__(self, o: Any) -> Any:
        return self.force() >= (o.force() if isinstance(o, LazyScalar) else o)

    def __eq__(self, o: Any) -> Any:  # type: ignore[override]
        return self.force() == (o.force() if isinstance(o, LazyScalar) else o)

    def __ne__(self, o: Any) -> Any:  # type: ignore[override]
        return self.force() != (o.force() if isinstance(o, LazyScalar) else o)

    # numpy ufuncs / functions on a LazyScalar: evaluate, then let numpy do its thing
    def __array_ufunc__(self, ufunc: Any, method: str, *inputs: Any, **kwargs: Any) -> Any:
        if method == "__call__" and ufunc in (np.add, np.subtract, np.multiply, np.true_divide, np.negative) and not kwargs:
            if ufunc is np.negative:
                return -self
            a, b = inputs
            if a is self:
                return {np.add: self.__add__, np.subtract: self.__sub__, np.multiply: self.__mul__, np.true_divide: self.__truediv__}[ufunc](b)
            return {np.add: self.__radd__, np.subtract: self.__rsub__, np.multiply: self.__rmul__, np.true_divide: self.__rtruediv__}[ufunc](a)
        ins = tuple(x.force() if isinstance(x, LazyScalar) else x for x in inputs)
        return getattr(ufunc, method)(*ins, **kwargs)


def force(x: Any) -> Any:
    """The plain value of ``x`` (recursively for lists / tuples / dicts)."""
    if isinstance(x, LazyScalar):
        return x.force()
    if isinstance(x, (list, tuple)):
        return type(x)(force(v) for v in x)
    if isinstance(x, dict):
        return {k: force(v) for k, v in x.items()}
    return x


_ = Dict  # typing re-export guard
