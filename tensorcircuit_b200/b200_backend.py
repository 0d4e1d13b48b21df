"""The backend object (``tc.backend`` / ``K``) for the B200 engine.

Mirrors the slice of the reference's ``ExtendedBackend`` method table that hot-path scripts
touch (tensorcircuit/backends/abstract_backend.py): small-tensor helpers (host numpy, they
only ever see parameters, expectation values and samples), ``jit`` (:1660-1683), ``vmap``
(:1685-1704), ``probability_sample`` (:1124-1157) and the random API (:914-1122).

O(2^n) data never goes through these helpers: states live in the engine and are only touched
by tcb200 kernels.  ``vmap`` calls the function once with :class:`BatchArray` arguments and the
engine runs batched kernels; with several ranks (torchrun) the batch is sharded across GPUs."""

from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import cons
from .batching import BatchArray, batch_of, is_batched, unwrap

Tensor = Any


def _np(x: Any) -> Any:
    from .circuit import DeviceArray

    if isinstance(x, DeviceArray):
        return x.numpy()
    return x


class B200Backend:
    name = "b200"

    def __init__(self) -> None:
        self.g = np.random.default_rng()

    # -- dtype / conversion ---------------------------------------------------------------
    def convert_to_tensor(self, a: Any, dtype: Any = None) -> Any:
        if is_batched(a):
            return a.astype(dtype) if dtype is not None else a
        a = np.asarray(_np(a))
        return a.astype(dtype) if dtype is not None else a

    def cast(self, a: Any, dtype: str) -> Any:
        if is_batched(a):
            if a.dtype.kind == "c" and np.dtype(dtype).kind != "c":
                a = a.real
            return a.astype(dtype)
        a = np.asarray(_np(a))
        if a.dtype.kind == "c" and np.dtype(dtype).kind != "c":
            a = a.real
        return a.astype(dtype)

    def numpy(self, a: Any) -> np.ndarray:
        if is_batched(a):
            raise TypeError("cannot leave vmap with .numpy(); return the value from the vmapped function")
        return np.asarray(_np(a))

    def is_tensor(self, a: Any) -> bool:
        from .circuit import DeviceArray

        return isinstance(a, (np.ndarray, BatchArray, DeviceArray))

    def dtype(self, a: Any) -> str:
        return str(np.asarray(_np(a)).dtype) if not is_batched(a) else str(a.dtype)

    def shape_tuple(self, a: Any) -> Tuple[int, ...]:
        return tuple(a.shape) if hasattr(a, "shape") else tuple(np.shape(a))

    def size(self, a: Any) -> int:
        return int(np.prod(self.shape_tuple(a)))

    sizen = size

    def i(self, dtype: Any = None) -> Any:
        return np.array(1j, dtype=dtype or cons.dtypestr)

    # -- creation -----------------------------------------------------------------------------
    def ones(self, shape: Sequence[int], dtype: Optional[str] = None) -> Any:
        return np.ones(shape, dtype=dtype or cons.dtypestr)

    def zeros(self, shape: Sequence[int], dtype: Optional[str] = None) -> Any:
        return np.zeros(shape, dtype=dtype or cons.dtypestr)

    def eye(self, N: int, dtype: Optional[str] = None, M: Optional[int] = None) -> Any:
        return np.eye(N, M, dtype=dtype or cons.dtypestr)

    def arange(self, start: int, stop: Optional[int] = None, step: int = 1) -> Any:
        return np.arange(start, stop, step) if stop is not None else np.arange(start)

    def onehot(self, a: Any, num: int) -> Any:
        return np.eye(num)[np.asarray(a)]

    one_hot = onehot

    # -- elementwise / linear algebra on small tensors ------------------------------------------------
    def _u(fn: Callable[..., Any]):  # type: ignore[misc]
        def f(self, *a: Any, **k: Any) -> Any:
            return fn(*[_np(x) for x in a], **k)

        return f

    sin = _u(np.sin)
    cos = _u(np.cos)
    tan = _u(np.tan)
    exp = _u(np.exp)
    log = _u(np.log)
    sqrt = _u(np.sqrt)
    abs = _u(np.abs)
    real = _u(np.real)
    imag = _u(np.imag)
    conj = _u(np.conj)
    kron = _u(np.kron)
    matmul = _u(np.matmul)
    tanh = _u(np.tanh)
    sign = _u(np.sign)
    mod = _u(np.mod)
    right_shift = _u(np.right_shift)
    left_shift = _u(np.left_shift)
    del _u

    def sum(self, a: Any, axis: Any = None, keepdims: bool = False) -> Any:
        a = _np(a)
        return np.sum(a, axis=axis) if not keepdims else np.sum(a, axis=axis, keepdims=True)

    def mean(self, a: Any, axis: Any = None, keepdims: bool = False) -> Any:
        return np.mean(_np(a), axis=axis)

    def stack(self, a: Sequence[Any], axis: int = 0) -> Any:
        return np.stack([_np(x) for x in a], axis=axis)

    def concat(self, a: Sequence[Any], axis: int = 0) -> Any:
        return np.concatenate([_np(x) for x in a], axis=axis)

    def reshape(self, a: Any, shape: Sequence[int]) -> Any:
        a = _np(a)
        return a.reshape(shape) if is_batched(a) else np.reshape(a, shape)

    def reshape2(self, a: Any) -> Any:
        from .gates import reshape2

        return reshape2(_np(a))

    def reshapem(self, a: Any) -> Any:
        from .gates import reshapem

        return reshapem(_np(a))

    def transpose(self, a: Any, perm: Optional[Sequence[int]] = None) -> Any:
        return np.transpose(_np(a), perm)

    # -- sparse operators (abstract_backend.py: coo_sparse_matrix / sparse_dense_matmul / to_dense / is_sparse) --
    def coo_sparse_matrix(self, indices: Any, values: Any, shape: Any) -> Any:
        """numpy_backend.py:303-309 returns a scipy COO matrix; so does this (host object -- it is uploaded
        once, on first use by ``operator_expectation`` / ``sparse_expectation``)."""
        import scipy.sparse as sp

        idx = np.asarray(indices)
        return sp.coo_matrix((np.asarray(values), (idx[:, 0], idx[:, 1])), shape=tuple(shape))

    def coo_sparse_matrix_from_numpy(self, a: Any) -> Any:
        return a

    def is_sparse(self, a: Any) -> bool:
        from .engine import DeviceCOO
        from .quantum import PauliSum

        return isinstance(a, (PauliSum, DeviceCOO)) or hasattr(a, "tocoo")

    def to_dense(self, sp_a: Any) -> Any:
        from .quantum import PauliSum

        if isinstance(sp_a, PauliSum):
            return sp_a.todense()
        return np.asarray(sp_a.todense())

    def sparse_dense_matmul(self, sp_a: Any, b: Any) -> Any:
        """host-side (scipy) product for small operands; the engine's own use of a sparse Hamiltonian is
        ``templates.measurements.sparse_expectation``, which never forms H|psi> on the host"""
        from .quantum import PauliSum

        if isinstance(sp_a, PauliSum):
            sp_a = sp_a.tocoo()
        return sp_a @ self.numpy(b)

    def adjoint(self, a: Any) -> Any:
        return np.conj(np.transpose(_np(a)))

    def expm(self, a: Any) -> Any:
        import scipy.linalg

        return scipy.linalg.expm(np.asarray(_np(a)))

    def tensordot(self, a: Any, b: Any, axes: Any = 2) -> Any:
        return np.tensordot(_np(a), _np(b), axes)

    def reverse(self, a: Any) -> Any:
        return np.asarray(_np(a))[::-1]

    def cumsum(self, a: Any, axis: Optional[int] = None) -> Any:
        return np.cumsum(_np(a), axis)

    def searchsorted(self, a: Any, v: Any, side: str = "left") -> Any:
        return np.searchsorted(np.asarray(_np(a)), np.asarray(_np(v)), side=side)

    def gather1d(self, operand: Any, indices: Any) -> Any:
        return np.asarray(_np(operand))[np.asarray(indices)]

    def unique_with_counts(self, a: Any, **kws: Any) -> Tuple[Any, Any]:
        return np.unique(np.asarray(a), return_counts=True)

    def scatter(self, operand: Any, indices: Any, updates: Any) -> Any:
        out = np.array(operand)
        out[tuple(np.asarray(indices).T)] = updates
        return out

    def argmax(self, a: Any, axis: int = 0) -> Any:
        return np.argmax(_np(a), axis=axis)

    def max(self, a: Any, axis: Any = None) -> Any:
        return np.max(_np(a), axis=axis)

    def min(self, a: Any, axis: Any = None) -> Any:
        return np.min(_np(a), axis=axis)

    def norm(self, a: Any) -> Any:
        return np.linalg.norm(np.asarray(_np(a)))

    # -- the rest of the small-tensor method table (abstract_backend.py:18-913, numpy_backend.py) --------
    # Host glue for parameter-sized data, so that scripts written against the reference's backend
    # object find the names they use; nothing here touches O(2^n) data.
    def _u1(fn: Callable[..., Any]):  # type: ignore[misc]
        def f(self, *a: Any, **k: Any) -> Any:
            return fn(*[_np(x) for x in a], **k)

        return f

    acos = _u1(np.arccos)
    acosh = _u1(np.arccosh)
    asin = _u1(np.arcsin)
    asinh = _u1(np.arcsinh)
    atan = _u1(np.arctan)
    atan2 = _u1(np.arctan2)
    atanh = _u1(np.arctanh)
    cosh = _u1(np.cosh)
    sinh = _u1(np.sinh)
    del _u1

    def relu(self, a: Any) -> Any:
        return np.maximum(_np(a), 0)

    def sigmoid(self, a: Any) -> Any:
        return 1.0 / (1.0 + np.exp(-_np(a)))

    def softmax(self, a: Sequence[Any], axis: Optional[int] = None) -> Any:
        a = np.asarray(_np(a))
        e = np.exp(a - np.max(a, axis=axis, keepdims=True))
        return e / np.sum(e, axis=axis, keepdims=True)

    def std(self, a: Any, axis: Any = None, keepdims: bool = False) -> Any:
        return np.std(np.asarray(_np(a)), axis=axis, keepdims=keepdims)

    def tile(self, a: Any, rep: Any) -> Any:
        return np.tile(np.asarray(_np(a)), np.asarray(rep))

    def argmin(self, a: Any, axis: int = 0) -> Any:
        return np.argmin(np.asarray(_np(a)), axis=axis)

    def copy(self, a: Any) -> Any:
        return np.array(np.asarray(_np(a)), copy=True)

    def stop_gradient(self, a: Any) -> Any:
        return a

    def cond(self, pred: bool, true_fun: Callable[[], Any], false_fun: Callable[[], Any]) -> Any:
        return true_fun() if pred else false_fun()

    def switch(self, index: Any, branches: Sequence[Callable[[], Any]]) -> Any:
        return branches[int(np.asarray(index))]()

    def scan(self, f: Callable[[Any, Any], Any], xs: Any, init: Any) -> Any:
        carry = init
        for x in np.asarray(_np(xs)):
            carry = f(carry, x)
        return carry

    def eigvalsh(self, a: Any) -> Any:
        return np.linalg.eigvalsh(np.asarray(_np(a)))

    def det(self, a: Any) -> Any:
        return np.linalg.det(np.asarray(_np(a)))

    def solve(self, A: Any, b: Any, **kws: Any) -> Any:
        return np.linalg.solve(np.asarray(_np(A)), np.asarray(_np(b)))

    def sqrtmh(self, a: Any) -> Any:
        e, v = np.linalg.eigh(np.asarray(_np(a)))
        return v @ np.diag(np.sqrt(e.astype(complex))) @ v.conj().T

    def tree_map(self, f: Callable[..., Any], *pytrees: Any) -> Any:
        t0 = pytrees[0]
        if isinstance(t0, (list, tuple)):
            return type(t0)(self.tree_map(f, *[p[i] for p in pytrees]) for i in range(len(t0)))
        if isinstance(t0, dict):
            return {k: self.tree_map(f, *[p[k] for p in pytrees]) for k in t0}
        return f(*pytrees)

    def tree_flatten(self, pytree: Any) -> Tuple[List[Any], Any]:
        leaves: List[Any] = []

        def go(t: Any) -> Any:
            if isinstance(t, (list, tuple)):
                return (type(t), [go(x) for x in t])
            if isinstance(t, dict):
                return (dict, {k: go(v) for k, v in t.items()})
            leaves.append(t)
            return None

        return leaves, go(pytree)

    def tree_unflatten(self, treedef: Any, leaves: Sequence[Any]) -> Any:
        it = iter(leaves)

        def go(d: Any) -> Any:
            if d is None:
                return next(it)
            kind, sub = d
            if kind is dict:
                return {k: go(v) for k, v in sub.items()}
            return kind(go(x) for x in sub)

        return go(treedef)

    def device(self, a: Any) -> str:
        from .circuit import DeviceArray

        return str(a.t.device) if isinstance(a, DeviceArray) else "cpu"

    def device_move(self, a: Any, dev: Any) -> Any:
        return a

    def implicit_randc(self, a: Any, shape: Union[int, Sequence[int]], p: Optional[Any] = None) -> Any:
        return self.stateful_randc(self.g, a, shape, p)

    def stateful_randc(self, g: Any, a: Any, shape: Union[int, Sequence[int]], p: Optional[Any] = None) -> Any:
        if isinstance(shape, int):
            shape = (shape,)
        if g is None:
            g = self.g
        pv = None if p is None else np.asarray(_np(p), dtype=np.float64)
        return g.choice(np.asarray(_np(a)) if not isinstance(a, int) else a, size=shape, p=None if pv is None else pv / pv.sum())

    # -- randomness (abstract_backend.py:914-1122; numpy_backend.py:245-308) -------------------
    def set_random_state(self, seed: Optional[Any] = None, get_only: bool = False) -> Any:
        g = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        if not get_only:
            self.g = g
        return g

    def get_random_state(self, seed: Optional[int] = None) -> Any:
        return self.set_random_state(seed, True)

    def random_split(self, key: Any) -> Tuple[Any, Any]:
        return key, key

    def implicit_randu(self, shape: Union[int, Sequence[int]] = 1, low: float = 0, high: float = 1, dtype: str = "32") -> Any:
        return self.stateful_randu(self.g, shape, low, high, dtype)

    def implicit_randn(self, shape: Union[int, Sequence[int]] = 1, mean: float = 0, stddev: float = 1, dtype: str = "32") -> Any:
        return self.stateful_randn(self.g, shape, mean, stddev, dtype)

    def stateful_randu(self, g: Any, shape: Union[int, Sequence[int]] = 1, low: float = 0, high: float = 1, dtype: str = "32") -> Any:
        if isinstance(shape, int):
            shape = (shape,)
        if g is None:
            g = self.g
        r = g.uniform(low=low, high=high, size=shape)
        return r.astype("float32" if dtype == "32" else "float64" if dtype == "64" else dtype)

    def stateful_randn(self, g: Any, shape: Union[int, Sequence[int]] = 1, mean: float = 0, stddev: float = 1, dtype: str = "32") -> Any:
        if isinstance(shape, int):
            shape = (shape,)
        if g is None:
            g = self.g
        r = g.normal(loc=mean, scale=stddev, size=shape)
        return r.astype("float32" if dtype == "32" else "float64" if dtype == "64" else dtype)

    def probability_sample(self, shots: int, p: Any, status: Optional[Any] = None, g: Any = None) -> Any:
        """abstract_backend.py:1124-1157 on the device: the probability vector becomes an
        amplitude vector sqrt(p) and goes through the engine's two-level CDF sampler."""
        from .circuit import DeviceArray
        from .engine import DeviceState
        import torch

        if status is None:
            status = self.stateful_randu(g, shape=[shots]) if g is not None else self.implicit_randu(shape=[shots])
        pv = p.t if isinstance(p, DeviceArray) else torch.from_numpy(np.ascontiguousarray(np.asarray(p, dtype=np.float64)))
        n = int(round(np.log2(pv.numel())))
        if 2**n != pv.numel():
            raise ValueError("probability_sample on the B200 backend needs a power-of-two length")
        st = DeviceState(n, "complex128", 1)
        st.load(pv)                 # (p, 0) as a state-shaped buffer ...
        st.sqrt_real_inplace()      # ... -> (sqrt(p), 0) by tcb200_probability_state: no torch op on 2^n data
        return st.sample(np.asarray(status, dtype=np.float64))

    # -- program transforms ------------------------------------------------------------------
    def jit(self, f: Callable[..., Any], static_argnums: Any = None, jit_compile: Any = None, **kws: Any) -> Callable[..., Any]:
        """Identity: kernels are precompiled and the fusion plan is cached by circuit structure
        (fusion.plan_structure), so there is nothing to trace -- and no staging time."""
        return f

    def vmap(self, f: Callable[..., Any], vectorized_argnums: Union[int, Sequence[int]] = 0) -> Callable[..., Any]:
        if isinstance(vectorized_argnums, int):
            vectorized_argnums = (vectorized_argnums,)
        vargs = tuple(vectorized_argnums)

        def wrapper(*args: Any, **kws: Any) -> Any:
            from .parallel import shard_batch, gather_batch

            args = list(args)
            B = None
            for i in vargs:
                a = np.asarray(_np(args[i]))
                if B is None:
                    B = a.shape[0]
                elif B != a.shape[0]:
                    raise ValueError("vectorised arguments disagree on the batch size")
                args[i] = a
            lo, hi = shard_batch(B)
            for i in vargs:
                args[i] = BatchArray(args[i][lo:hi])
            out = f(*args, **kws)
            out = unwrap(out, hi - lo)
            return gather_batch(out, B)

        return wrapper

    # Gradients of losses built from expectation values (autodiff.py): an exact shift rule on the
    # gate matrices, all shifted circuits evaluated as one batch on the same kernels.
    def value_and_grad(self, f: Callable[..., Any], argnums: Any = 0, has_aux: bool = False) -> Callable[..., Any]:
        from . import autodiff

        return autodiff.value_and_grad(f, argnums=argnums, has_aux=has_aux)

    def grad(self, f: Callable[..., Any], argnums: Any = 0, has_aux: bool = False) -> Callable[..., Any]:
        from . import autodiff

        return autodiff.grad(f, argnums=argnums, has_aux=has_aux)

    def vectorized_value_and_grad(self, f: Callable[..., Any], argnums: Any = 0, vectorized_argnums: Any = 0, has_aux: bool = False) -> Callable[..., Any]:
        from . import autodiff

        return autodiff.vectorized_value_and_grad(f, argnums=argnums, vectorized_argnums=vectorized_argnums, has_aux=has_aux)

    vvag = vectorized_value_and_grad

    # vjp / jvp / Jacobians / Hessian on top of the adjoint sweep (abstract_backend.py:1461-1658)
    def vjp(self, f: Callable[..., Any], inputs: Any, v: Any) -> Any:
        from . import autodiff

        return autodiff.vjp(f, inputs, v)

    def jvp(self, f: Callable[..., Any], inputs: Any, v: Any) -> Any:
        from . import autodiff

        return autodiff.jvp(f, inputs, v)

    def jacrev(self, f: Callable[..., Any], argnums: Any = 0) -> Callable[..., Any]:
        from . import autodiff

        return autodiff.jacrev(f, argnums=argnums)

    def jacfwd(self, f: Callable[..., Any], argnums: Any = 0) -> Callable[..., Any]:
        from . import autodiff

        return autodiff.jacfwd(f, argnums=argnums)

    def hessian(self, f: Callable[..., Any], argnums: Any = 0) -> Callable[..., Any]:
        from . import autodiff

        return autodiff.hessian(f, argnums=argnums)


_INSTANCE: Optional[B200Backend] = None


def get_backend(name: Optional[str] = None) -> B200Backend:
    global _INSTANCE
    if _INSTANCE is None:
        _INSTANCE = B200Backend()
    return _INSTANCE


cons.backend = get_backend()
