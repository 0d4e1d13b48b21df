"""Measurement shortcuts (tensorcircuit/templates/measurements.py:18-208, 290-335) mapped onto the
multi-term Pauli expectation kernels and, for generic sparse / dense operators, onto the COO
expectation kernel (csrc/sparse.cu)."""

from __future__ import annotations

from typing import Any, Optional, Sequence

import numpy as np

from .. import cons
from ..batching import BatchArray, is_batched
from ..circuit import Circuit

Tensor = Any


def any_measurements(c: Circuit, structures: Tensor, onehot: bool = False, reuse: bool = True) -> Tensor:
    """measurements.py:18-87: expectation of the Pauli string given as a *tensor* (so it can be
    vmapped): ``structures`` is [n] ints in 0..3 (``onehot=True``) or [n, 4] weights."""
    n = c._nqubits
    if onehot:
        if is_batched(structures):
            raise NotImplementedError("batched integer structures: pass one-hot weights [n, 4]")
        ps = [int(v) for v in np.asarray(structures).reshape(-1)]
        return np.real(c.expectation_ps(ps=ps))
    s = structures
    obs = []
    from .. import gates

    paulis = [np.eye(2), gates._x_matrix, gates._y_matrix, gates._z_matrix]
    for i in range(n):
        op = sum(s[i, k] * paulis[k] for k in range(4))
        obs.append([gates.Gate(op), (i,)])
    return np.real(c.expectation(*obs, reuse=reuse))


parameterized_measurements = any_measurements


def any_local_measurements(c: Circuit, structures: Tensor, onehot: bool = False, reuse: bool = True) -> Tensor:
    """measurements.py:90-153: vector of single-site expectations, one per qubit."""
    n = c._nqubits
    ps_all = []
    st = [int(v) for v in np.asarray(structures).reshape(-1)] if onehot else None
    if st is None:
        raise NotImplementedError("weighted local measurements: pass onehot=True")
    for i in range(n):
        ps = [0] * n
        ps[i] = st[i]
        ps_all.append(ps)
    return np.real(c.expectation_ps_many(ps_all))


parameterized_local_measurements = any_local_measurements


def pauli_sum_expectation(c: Circuit, pss: Sequence[Sequence[int]], weights: Optional[Sequence[float]] = None) -> Tensor:
    """sum_t w_t <P_t> for a Hamiltonian given as Pauli strings -- all terms in as few reads of
    the state as possible (the one-pass counterpart of looping over ``expectation_ps``)."""
    vals = c.expectation_ps_many(pss)
    w = np.ones(len(pss)) if weights is None else np.asarray(weights)
    if is_batched(vals):
        return BatchArray(np.real(vals.a) @ w)
    return np.real(np.sum(w * vals))


def spin_glass_measurements(c: Circuit, g: Any, reuse: bool = True) -> Tensor:
    """measurements.py:290-335: sum_ij w_ij <Z_i Z_j> + sum_i h_i <Z_i>."""
    n = c._nqubits
    pss, ws = [], []
    for e1, e2 in g.edges:
        ps = [0] * n
        ps[e1] = ps[e2] = 3
        pss.append(ps)
        ws.append(g[e1][e2].get("weight", 1.0))
    for i in g.nodes:
        w = g.nodes[i].get("weight", 0.0)
        if w != 0.0:
            ps = [0] * n
            ps[i] = 3
            pss.append(ps)
            ws.append(w)
    return pauli_sum_expectation(c, pss, ws)


def heisenberg_measurements(c: Circuit, g: Any, hzz: float = 1.0, hxx: float = 1.0, hyy: float = 1.0, hz: float = 0.0,
                            hx: float = 0.0, hy: float = 0.0, reuse: bool = True) -> Tensor:
    """measurements.py:211-287:  sum_e w_e (hxx XX + hyy YY + hzz ZZ) + sum_v (hx X + hy Y + hz Z).
    All strings go through the multi-term kernels in one call (the ZZ / Z ones share a single
    streaming read of the state)."""
    n = c._nqubits
    pss, ws = [], []
    for e1, e2 in g.edges:
        w = g[e1][e2].get("weight", 1.0)
        for p, h in ((3, hzz), (2, hyy), (1, hxx)):
            ps = [0] * n
            ps[e1] = ps[e2] = p
            pss.append(ps)
            ws.append(w * h)
    for p, h in ((1, hx), (2, hy), (3, hz)):
        if h != 0:
            for i in range(len(g.nodes)):
                ps = [0] * n
                ps[i] = p
                pss.append(ps)
                ws.append(h)
    return pauli_sum_expectation(c, pss, ws)


def _operator_on_device(hamiltonian: Any) -> Any:
    """scipy sparse / dense ndarray -> engine.DeviceCOO, cached on scipy objects (a VQE loop passes the
    same Hamiltonian every iteration: it is uploaded once)."""
    from ..engine import DeviceCOO

    if isinstance(hamiltonian, DeviceCOO):
        return hamiltonian
    cached = getattr(hamiltonian, "_tcb200_device", None)
    if cached is not None:
        return cached
    if hasattr(hamiltonian, "tocoo"):
        op = DeviceCOO.from_scipy(hamiltonian)
        try:
            hamiltonian._tcb200_device = op
        except AttributeError:
            pass
        return op
    a = np.asarray(hamiltonian)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("operator_expectation needs a square matrix")
    r, c = np.nonzero(a)
    return DeviceCOO(r, c, a[r, c], a.shape[0])


def sparse_expectation(c: Circuit, hamiltonian: Tensor) -> Tensor:
    """measurements.py:173-188.  A :class:`quantum.PauliSum` (what ``PauliStringSum2COO`` returns) goes
    through the Pauli kernels string by string -- no matrix; a scipy sparse matrix / ``DeviceCOO`` /
    dense array goes through the COO expectation kernel with the operator resident on the device."""
    from ..quantum import PauliSum

    if isinstance(hamiltonian, PauliSum):
        if hamiltonian.nqubits != c._nqubits:
            raise ValueError("Hamiltonian on %d qubits, circuit on %d" % (hamiltonian.nqubits, c._nqubits))
        return pauli_sum_expectation(c, hamiltonian.ls.tolist(), hamiltonian.weight)
    op = _operator_on_device(hamiltonian)
    st = c._ensure_state()
    r = np.real(st.coo_expectation(op))
    rd = np.float32 if c._dtype == "complex64" else np.float64
    if c._batch is None:
        return rd(r[0])
    return BatchArray(r.astype(rd))


def mpo_expectation(c: Circuit, mpo: Any) -> Tensor:
    """measurements.py:191-208 contracts a tensornetwork ``QuOperator`` with the circuit's node graph;
    both live in the tensor-network layer that this engine replaces (no node graph exists here)."""
    raise NotImplementedError("mpo_expectation needs the tensornetwork QuOperator layer, which is outside the statevector hot path; "
                              "pass the Hamiltonian as Pauli strings (quantum.PauliStringSum2COO) or as a sparse matrix")


def operator_expectation(c: Circuit, hamiltonian: Any) -> Tensor:
    """measurements.py:156-170: dense matrix, sparse matrix or Pauli sum (MPO: see mpo_expectation)."""
    if type(hamiltonian).__name__ == "QuOperator":
        return mpo_expectation(c, hamiltonian)
    return sparse_expectation(c, hamiltonian)
