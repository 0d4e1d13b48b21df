"""Measurement post-processing and Pauli-string utilities used on the hot path (host side,
O(shots) or O(n) data).  Mirrors tensorcircuit/quantum.py:1025-1068 (ps2xyz/xyz2ps),
:2048-2157 (count conversions), :2217-2368 (measurement_counts / sample2all) and :2371-2451
(spin_by_basis / correlations)."""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

from functools import partial

import numpy as np

Tensor = Any


def ps2xyz(ps: List[int]) -> Dict[str, List[int]]:
    xyz: Dict[str, List[int]] = {"x": [], "y": [], "z": []}
    for i, j in enumerate(ps):
        j = int(j)
        if j == 1:
            xyz["x"].append(i)
        if j == 2:
            xyz["y"].append(i)
        if j == 3:
            xyz["z"].append(i)
    return xyz


def xyz2ps(xyz: Dict[str, List[int]], n: Optional[int] = None) -> List[int]:
    if n is None:
        n = max(xyz.get("x", []) + xyz.get("y", []) + xyz.get("z", [])) + 1
    ps = [0 for _ in range(n)]
    for i in range(n):
        if i in xyz.get("x", []):
            ps[i] = 1
        elif i in xyz.get("y", []):
            ps[i] = 2
        elif i in xyz.get("z", []):
            ps[i] = 3
    return ps


def sample_int2bin(sample: Tensor, n: int) -> Tensor:
    sample = np.asarray(sample)
    return np.mod(np.right_shift(sample[..., None], np.arange(n)[::-1]), 2)


def sample_bin2int(sample: Tensor, n: int) -> Tensor:
    power = np.array([2**j for j in reversed(range(n))])
    return np.sum(np.asarray(sample) * power, axis=-1)


def sample2count(sample: Tensor, n: int, jittable: bool = True) -> Tuple[Tensor, Tensor]:
    return np.unique(np.asarray(sample), return_counts=True)


def count_s2d(srepr: Tuple[Tensor, Tensor], n: int) -> Tensor:
    out = np.zeros([2**n], dtype=np.asarray(srepr[1]).dtype)
    out[np.asarray(srepr[0]).reshape(-1)] = srepr[1]
    return out


counts_v2t = count_s2d


def count_d2s(drepr: Tensor, eps: float = 1e-7) -> Tuple[Tensor, Tensor]:
    drepr = np.asarray(drepr)
    x = np.nonzero(np.abs(drepr) > eps)[0]
    return x, drepr[x]


count_t2v = count_d2s


def count_vector2dict(count: Tensor, n: int, key: str = "bin") -> Dict[Any, int]:
    count = np.asarray(count)
    d = {i: count[i].item() for i in range(2**n)}
    if key == "int":
        return d
    return {bin(k)[2:].zfill(n): v for k, v in d.items()}


def count_tuple2dict(count: Tuple[Tensor, Tensor], n: int, key: str = "bin") -> Dict[Any, int]:
    d = {int(i): int(j) for i, j in zip(count[0], count[1]) if i >= 0}
    if key == "int":
        return d
    return {bin(k)[2:].zfill(n): v for k, v in d.items()}


def sample2all(sample: Tensor, n: int, format: str = "count_vector", jittable: bool = False, format_: Optional[str] = None) -> Any:
    if format_ is not None:
        format = format_
    sample = np.asarray(sample)
    if sample.ndim not in (1, 2):
        raise ValueError("unrecognized tensor shape for sample")
    # each representation is built only when the requested format needs it: the [shots, n] bit array of 10^6 shots
    # at n = 34 is 272 MB and ~170 ms of host time, which "sample_int" never looks at
    if format == "sample_int":
        return sample if sample.ndim == 1 else sample_bin2int(sample, n)
    if format == "sample_bin":
        return sample_int2bin(sample, n) if sample.ndim == 1 else sample
    sample_int = sample if sample.ndim == 1 else sample_bin2int(sample, n)
    count_tuple = sample2count(sample_int, n, jittable)
    if format == "count_tuple":
        return count_tuple
    if format == "count_vector":
        return count_s2d(count_tuple, n)
    if format == "count_dict_bin":
        return count_tuple2dict(count_tuple, n, key="bin")
    if format == "count_dict_int":
        return count_tuple2dict(count_tuple, n, key="int")
    raise ValueError("unsupported format %s for finite shots measurement" % format)


def measurement_counts(
    state: Tensor,
    counts: Optional[int] = 8192,
    format: str = "count_vector",
    is_prob: bool = False,
    random_generator: Optional[Any] = None,
    status: Optional[Tensor] = None,
    jittable: bool = False,
    format_: Optional[str] = None,
) -> Any:
    """quantum.py:2217-2318: simulate ``counts`` shots on a state (or probability) vector.
    The state goes to the device and through the engine's CDF sampler."""
    from . import cons
    from .circuit import Circuit

    if format_ is not None:
        format = format_
    v = np.asarray(state)
    n = int(round(np.log2(v.shape[0])))
    if v.ndim == 2:  # density matrix: use its diagonal
        v = np.sqrt(np.abs(np.diagonal(v)))
    elif is_prob:
        v = np.sqrt(np.abs(v))
    if counts is None or counts <= 0:
        p = np.abs(v) ** 2
        p = p / np.sum(p)
        if counts is not None and counts < 0:
            p = p * (-counts)
        if format == "count_vector":
            return p
        if format == "count_tuple":
            return count_d2s(p)
        if format == "count_dict_bin":
            return count_vector2dict(p, n, key="bin")
        if format == "count_dict_int":
            return count_vector2dict(p, n, key="int")
        raise ValueError("unsupported format %s for analytical measurement" % format)
    c = Circuit(n, inputs=v)
    return c.sample(batch=counts, allow_state=True, format=format, random_generator=random_generator, status=status)


measurement_results = measurement_counts


def spin_by_basis(n: int, m: int, elements: Tuple[int, int] = (1, -1)) -> Tensor:
    r = np.arange(2**n)
    b = (r >> (n - 1 - m)) & 1
    return np.where(b == 0, elements[0], elements[1])


def correlation_from_samples(index: Sequence[int], results: Tensor, n: int) -> Tensor:
    results = np.asarray(results)
    if results.ndim == 1:
        results = sample_int2bin(results, n)
    results = 1 - results * 2
    r = results[:, index[0]]
    for i in index[1:]:
        r = r * results[:, i]
    return np.mean(r.astype(np.float64))


def correlation_from_counts(index: Sequence[int], results: Tensor) -> Tensor:
    results = np.asarray(results, dtype=np.float64)
    results = results / np.sum(results)
    n = int(round(np.log2(results.shape[0])))
    for i in index:
        results = results * spin_by_basis(n, i)
    return np.sum(results)


# ---- Pauli-sum Hamiltonians (quantum.py:1163-1482) ----------------------------------------------
class PauliSum:
    """A Hamiltonian  sum_t w_t P_t  kept as its Pauli strings.

    The reference turns such a sum into a COO matrix with nterms * 2^n stored elements
    (``PauliStringSum2COO``, quantum.py:1304-1376) and multiplies it with the state
    (``sparse_expectation``).  Here the object returned by ``PauliStringSum2COO`` keeps the strings:
    ``operator_expectation`` / ``sparse_expectation`` evaluate them with the multi-term Pauli kernels
    (a few reads of the state, no matrix), and the matrix is only materialised on request
    (``tocoo`` / ``todense``, host, scipy -- same element values as the reference's)."""

    def __init__(self, ls: Sequence[Sequence[int]], weight: Optional[Sequence[float]] = None):
        self.ls = np.real(np.asarray(ls)).astype(np.int64).reshape(len(ls), -1)  # (tc.array_to_tensor hands over complex arrays)
        if self.ls.size and (self.ls.min() < 0 or self.ls.max() > 3):
            raise ValueError("Pauli strings are sequences of 0 (I), 1 (X), 2 (Y), 3 (Z)")
        self.weight = np.ones(len(self.ls)) if weight is None else np.asarray(weight).reshape(-1)
        if len(self.weight) != len(self.ls):
            raise ValueError("one weight per Pauli string")
        self.nqubits = int(self.ls.shape[1])
        self.shape = (1 << self.nqubits, 1 << self.nqubits)

    def __len__(self) -> int:
        return len(self.ls)

    def tocoo(self) -> Any:
        return PauliStringSum2COO(self.ls, self.weight, numpy=True)

    def todense(self) -> Any:
        return np.asarray(self.tocoo().todense())


def ps2coo_core(idx_x: int, idx_y: int, idx_z: int, weight: Any, nqubits: int) -> Any:
    """quantum.py:1461-1482: element (r, r ^ x ^ y) = (1 - 2 parity(r & (y | z))) (-i)^{ny} w."""
    import scipy.sparse as sp

    s = 1 << nqubits
    idx1 = np.arange(s, dtype=np.int64)
    idx2 = idx1 ^ np.int64(idx_x) ^ np.int64(idx_y)
    tmp = idx1 & np.int64(idx_y | idx_z)
    e = np.zeros(s, dtype=np.int64)
    for i in range(nqubits):
        e ^= (tmp >> i) & 1
    ny = bin(int(idx_y)).count("1") % 4
    values = (1 - 2 * e) * ((-1.0j) ** ny) * weight
    return sp.coo_matrix((values.astype(dtypestr_()), (idx1, idx2)), shape=(s, s))


def dtypestr_() -> str:
    from . import cons

    return cons.dtypestr


def PauliString2COO(l: Sequence[int], weight: Optional[float] = None) -> Any:
    """quantum.py:1415-1458 (host, scipy)."""
    n = len(l)
    ix = iy = iz = 0
    for j, p in enumerate(l):
        b = 1 << (n - j - 1)
        if p == 1:
            ix |= b
        elif p == 2:
            iy |= b
        elif p == 3:
            iz |= b
    return ps2coo_core(ix, iy, iz, 1.0 if weight is None else weight, n)


def PauliStringSum2COO(ls: Sequence[Sequence[int]], weight: Optional[Sequence[float]] = None, numpy: bool = False) -> Any:
    """quantum.py:1304-1360.  ``numpy=True``: the scipy COO matrix, element for element the
    reference's.  Otherwise the backend-side operator, which here is a :class:`PauliSum` -- the
    strings themselves; ``operator_expectation`` / ``sparse_expectation`` / ``backend.is_sparse`` /
    ``backend.to_dense`` / ``backend.sparse_dense_matmul`` accept it wherever the reference takes its
    sparse tensor."""
    if not numpy:
        return PauliSum(ls, weight)
    ls = np.real(np.asarray(ls)).astype(np.int64)
    w = np.ones(len(ls)) if weight is None else np.asarray(weight)
    acc = None
    for i in range(len(ls)):
        m = PauliString2COO(ls[i], w[i]).tocsr()
        acc = m if acc is None else acc + m
    return acc.tocoo()


PauliStringSum2COO_numpy = partial(PauliStringSum2COO, numpy=True)
PauliStringSum2COO_tf = PauliStringSum2COO


def PauliStringSum2Dense(ls: Sequence[Sequence[int]], weight: Optional[Sequence[float]] = None, numpy: bool = False) -> Any:
    """quantum.py:1257-1284."""
    return np.asarray(PauliStringSum2COO(ls, weight, numpy=True).todense())


def heisenberg_hamiltonian(g: Any, hzz: float = 1.0, hxx: float = 1.0, hyy: float = 1.0, hz: float = 0.0, hx: float = 0.0,
                           hy: float = 0.0, sparse: bool = True, numpy: bool = False) -> Any:
    """quantum.py:1163-1254: same string order as the reference (per edge zz, xx, yy; per node z, x, y)."""
    n = len(g.nodes)
    ls, weight = [], []
    for e in g.edges:
        for p, h in ((3, hzz), (1, hxx), (2, hyy)):
            if h != 0:
                r = [0] * n
                r[e[0]] = r[e[1]] = p
                ls.append(r)
                weight.append(h)
    for node in g.nodes:
        for p, h in ((3, hz), (1, hx), (2, hy)):
            if h != 0:
                r = [0] * n
                r[node] = p
                ls.append(r)
                weight.append(h)
    if sparse:
        return PauliStringSum2COO(ls, weight, numpy=numpy)
    return PauliStringSum2Dense(ls, weight)
