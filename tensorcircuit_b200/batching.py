"""Batched host values for ``backend.vmap``.

The reference's ``K.vmap(f)`` hands *batched parameters* to an unmodified ``f`` that builds a
Circuit (tensorcircuit/backends/jax_backend.py:718-730 traces ``f`` once with batch tracers).
Here ``f`` is likewise called ONCE: vectorised arguments are wrapped in :class:`BatchArray`,
whose leading axis is a hidden batch axis.  Gate formulas (cos/sin/exp products of small
matrices) broadcast over it, the Circuit notices batched gate matrices and runs the batched
CUDA kernels on a ``[B, 2^n]`` state, and every query returns a :class:`BatchArray` again.
Only parameter-sized data lives here (host, float64); nothing O(2^n)."""

from __future__ import annotations

import math
from typing import Any, Callable, Dict, Sequence, Tuple

import numpy as np

_HANDLED: Dict[Callable[..., Any], Callable[..., Any]] = {}


def _implements(np_func: Callable[..., Any]) -> Callable[..., Any]:
    def deco(f: Callable[..., Any]) -> Callable[..., Any]:
        _HANDLED[np_func] = f
        return f

    return deco


class BatchArray:
    """``a`` has shape ``[B, *shape]``; ``shape`` is what user code sees."""

    __array_priority__ = 1000

    def __init__(self, data: Any):
        self.a = np.asarray(data)
        assert self.a.ndim >= 1

    # -- introspection -----------------------------------------------------------------
    @property
    def batch(self) -> int:
        return int(self.a.shape[0])

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.a.shape[1:])

    @property
    def ndim(self) -> int:
        return self.a.ndim - 1

    @property
    def dtype(self) -> Any:
        return self.a.dtype

    @property
    def size(self) -> int:
        return math.prod(self.a.shape[1:])

    def __len__(self) -> int:
        if self.ndim == 0:
            raise TypeError("len() of a batched scalar")
        return self.shape[0]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __repr__(self) -> str:
        return "BatchArray(batch=%d, shape=%s, dtype=%s)" % (self.batch, self.shape, self.dtype)

    def __float__(self) -> float:
        raise TypeError("a batched (vmap) value has no single float value; reduce it inside vmap")

    __complex__ = __float__
    __int__ = __float__
    __bool__ = __float__

    # -- structure ---------------------------------------------------------------------
    def __getitem__(self, idx: Any) -> "BatchArray":
        if not isinstance(idx, tuple):
            idx = (idx,)
        return BatchArray(self.a[(slice(None),) + idx])

    def reshape(self, *shape: Any) -> "BatchArray":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return BatchArray(self.a.reshape((self.batch,) + tuple(shape)))

    def astype(self, dtype: Any) -> "BatchArray":
        return BatchArray(self.a.astype(dtype))

    def _aligned(self, nd: int) -> np.ndarray:
        return self.a.reshape((self.batch,) + (1,) * (nd - self.ndim) + self.shape)

    @property
    def real(self) -> "BatchArray":
        return BatchArray(self.a.real)

    @property
    def imag(self) -> "BatchArray":
        return BatchArray(self.a.imag)

    def conj(self) -> "BatchArray":
        return BatchArray(self.a.conj())

    def sum(self, axis: Any = None) -> "BatchArray":
        return _sum(self, axis=axis)

    def mean(self, axis: Any = None) -> "BatchArray":
        return _mean(self, axis=axis)

    # -- numpy protocols -----------------------------------------------------------------
    def __array_ufunc__(self, ufunc: Any, method: str, *inputs: Any, **kw: Any) -> Any:
        if method != "__call__" or kw.get("out") is not None:
            return NotImplemented
        nd = 0
        for x in inputs:
            nd = max(nd, x.ndim if isinstance(x, BatchArray) else np.ndim(x))
        conv = [x._aligned(nd) if isinstance(x, BatchArray) else np.asarray(x) for x in inputs]
        out = ufunc(*conv, **kw)
        if isinstance(out, tuple):
            return tuple(BatchArray(o) for o in out)
        return BatchArray(out)

    def __array_function__(self, func: Any, types: Any, args: Any, kwargs: Any) -> Any:
        if func not in _HANDLED:
            return NotImplemented
        return _HANDLED[func](*args, **kwargs)

    def __array__(self, dtype: Any = None, copy: Any = None) -> np.ndarray:
        raise TypeError("cannot convert a batched (vmap) value to a plain array inside vmap")

    # arithmetic through the ufunc protocol
    def __add__(self, o): return np.add(self, o)
    def __radd__(self, o): return np.add(o, self)
    def __sub__(self, o): return np.subtract(self, o)
    def __rsub__(self, o): return np.subtract(o, self)
    def __mul__(self, o): return np.multiply(self, o)
    def __rmul__(self, o): return np.multiply(o, self)
    def __truediv__(self, o): return np.true_divide(self, o)
    def __rtruediv__(self, o): return np.true_divide(o, self)
    def __pow__(self, o): return np.power(self, o)
    def __rpow__(self, o): return np.power(o, self)
    def __neg__(self): return np.negative(self)
    def __pos__(self): return self
    def __abs__(self): return np.absolute(self)

    def __matmul__(self, o):
        return _matmul(self, o)

    def __rmatmul__(self, o):
        return _matmul(o, self)


def is_batched(x: Any) -> bool:
    return isinstance(x, BatchArray)


def batch_of(*xs: Any) -> Any:
    b = None
    for x in xs:
        if isinstance(x, BatchArray):
            if b is not None and b != x.batch:
                raise ValueError("inconsistent vmap batch sizes %d vs %d" % (b, x.batch))
            b = x.batch
    return b


def _axis(ax: Any, nd: int) -> Any:
    """per-example axis -> axis of the underlying array"""
    if ax is None:
        return tuple(range(1, nd + 1))
    if isinstance(ax, (tuple, list)):
        return tuple(_axis(a, nd) for a in ax)
    return ax + 1 if ax >= 0 else ax


@_implements(np.sum)
def _sum(x: BatchArray, axis: Any = None, **kw: Any) -> BatchArray:
    return BatchArray(np.sum(x.a, axis=_axis(axis, x.ndim), **kw))


@_implements(np.mean)
def _mean(x: BatchArray, axis: Any = None, **kw: Any) -> BatchArray:
    return BatchArray(np.mean(x.a, axis=_axis(axis, x.ndim), **kw))


@_implements(np.real)
def _real(x: BatchArray) -> BatchArray:
    return x.real


@_implements(np.imag)
def _imag(x: BatchArray) -> BatchArray:
    return x.imag


@_implements(np.conj)
def _conj(x: BatchArray) -> BatchArray:
    return x.conj()


@_implements(np.reshape)
def _reshape(x: BatchArray, shape: Any, *a: Any, **k: Any) -> BatchArray:
    return x.reshape(shape)


@_implements(np.shape)
def _shape(x: BatchArray) -> Tuple[int, ...]:
    return x.shape


@_implements(np.ndim)
def _ndim(x: BatchArray) -> int:
    return x.ndim


def _lift(x: Any, b: int) -> np.ndarray:
    if isinstance(x, BatchArray):
        return x.a
    x = np.asarray(x)
    return np.broadcast_to(x, (b,) + x.shape)


@_implements(np.stack)
def _stack(xs: Sequence[Any], axis: int = 0, **kw: Any) -> BatchArray:
    b = batch_of(*xs)
    return BatchArray(np.stack([_lift(x, b) for x in xs], axis=_axis(axis, 0) if axis >= 0 else axis, **kw))


@_implements(np.concatenate)
def _concatenate(xs: Sequence[Any], axis: int = 0, **kw: Any) -> BatchArray:
    b = batch_of(*xs)
    return BatchArray(np.concatenate([_lift(x, b) for x in xs], axis=_axis(axis, 0) if axis >= 0 else axis, **kw))


@_implements(np.matmul)
def _matmul(x: Any, y: Any) -> BatchArray:
    b = batch_of(x, y)
    return BatchArray(np.matmul(_lift(x, b), _lift(y, b)))


@_implements(np.kron)
def _kron(x: Any, y: Any) -> BatchArray:
    b = batch_of(x, y)
    xa, ya = _lift(x, b), _lift(y, b)
    assert xa.ndim == 3 and ya.ndim == 3, "batched kron expects matrices"
    out = np.einsum("bij,bkl->bikjl", xa, ya)
    return BatchArray(out.reshape(b, xa.shape[1] * ya.shape[1], xa.shape[2] * ya.shape[2]))


@_implements(np.transpose)
def _transpose(x: BatchArray, axes: Any = None) -> BatchArray:
    if axes is None:
        axes = tuple(reversed(range(x.ndim)))
    return BatchArray(np.transpose(x.a, (0,) + tuple(a + 1 for a in axes)))


@_implements(np.where)
def _where(c: Any, x: Any, y: Any) -> BatchArray:
    return np.add(np.multiply(c, x), np.multiply(np.subtract(1, c), y))  # type: ignore


def unwrap(x: Any, batch: int) -> Any:
    """vmap output: stack on axis 0 (tensorcircuit/backends/numpy_backend.py:394-418)."""
    if isinstance(x, BatchArray):
        return x.a
    if getattr(x, "batched_on_device", False):  # circuit.BatchedDeviceArray: leading axis IS the batch
        return x
    if isinstance(x, (tuple, list)):
        return type(x)(unwrap(e, batch) for e in x)
    if isinstance(x, dict):
        return {k: unwrap(v, batch) for k, v in x.items()}
    x = np.asarray(x)
    return np.broadcast_to(x, (batch,) + x.shape).copy()
