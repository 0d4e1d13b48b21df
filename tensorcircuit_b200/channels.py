"""Kraus channels used by the Monte-Carlo trajectory methods of ``Circuit``
(``c.depolarizing / amplitudedamping / phasedamping / reset``, ``unitary_kraus``,
``general_kraus``): the Kraus-list definitions of tensorcircuit/channels.py:56-326.
Host side only (2x2 / 4^n matrices in complex128; parameters may be vmap-batched)."""

from __future__ import annotations

from typing import Any, List, Sequence

import numpy as np

from . import gates
from .gates import Gate

Tensor = Any


class KrausList(list):  # channels.py:19-29
    def __init__(self, iterable: Sequence[Gate], name: str, is_unitary: bool):
        super().__init__(iterable)
        self.name = name
        self.is_unitary = is_unitary


def _sqrt(a: Tensor) -> Tensor:
    """channels.py:32-46: sqrt in the complex domain so that negative inputs do not give nan"""
    return np.sqrt(gates.num_to_tensor(a))


_E00 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
_E01 = np.array([[0, 1], [0, 0]], dtype=np.complex128)
_E10 = np.array([[0, 0], [1, 0]], dtype=np.complex128)
_E11 = np.array([[0, 0], [0, 1]], dtype=np.complex128)


def depolarizingchannel(px: float, py: float, pz: float) -> Sequence[Gate]:  # channels.py:56-101
    i = Gate(_sqrt(1 - px - py - pz) * gates._i_matrix)
    x = Gate(_sqrt(px) * gates._x_matrix)
    y = Gate(_sqrt(py) * gates._y_matrix)
    z = Gate(_sqrt(pz) * gates._z_matrix)
    return KrausList([i, x, y, z], name="depolarizing", is_unitary=True)


def generaldepolarizingchannel(p: Any, num_qubits: int = 1) -> Sequence[Gate]:  # channels.py:140-214
    """p: scalar (same probability for all 4^n-1 non-identity Paulis) or a list of 4^n-1 values."""
    import itertools

    paulis = [gates._i_matrix, gates._x_matrix, gates._y_matrix, gates._z_matrix]
    strings = list(itertools.product(range(4), repeat=num_qubits))
    if np.ndim(p) == 0:
        probs = [1 - (4**num_qubits - 1) * p] + [p] * (4**num_qubits - 1)
    else:
        p = list(p)
        assert len(p) == 4**num_qubits - 1
        probs = [1 - sum(p)] + p
    ks = []
    for pr, s in zip(probs, strings):
        m = np.array([[1.0 + 0j]])
        for a in s:
            m = np.kron(m, paulis[a])
        ks.append(Gate(gates.reshape2(_sqrt(pr) * m)))
    return KrausList(ks, name="depolarizing", is_unitary=True)


def isotropicdepolarizingchannel(p: float, num_qubits: int = 1) -> Sequence[Gate]:  # channels.py:104-137
    return generaldepolarizingchannel(p / (4**num_qubits - 1), num_qubits)


def amplitudedampingchannel(gamma: float, p: float) -> Sequence[Gate]:  # channels.py:217-267
    m0 = Gate(_sqrt(p) * (_E00 + _sqrt(1 - gamma) * _E11))
    m1 = Gate(_sqrt(p) * (_sqrt(gamma) * _E01))
    m2 = Gate(_sqrt(1 - p) * (_sqrt(1 - gamma) * _E00 + _E11))
    m3 = Gate(_sqrt(1 - p) * (_sqrt(gamma) * _E10))
    return KrausList([m0, m1, m2, m3], name="amplitude_damping", is_unitary=False)


def resetchannel() -> Sequence[Gate]:  # channels.py:270-294
    return KrausList([Gate(_E00), Gate(_E01)], name="reset", is_unitary=False)


def phasedampingchannel(gamma: float) -> Sequence[Gate]:  # channels.py:297-325
    m0 = Gate(1.0 * (_E00 + _sqrt(1 - gamma) * _E11))
    m1 = Gate(_sqrt(gamma) * _E11)
    return KrausList([m0, m1], name="phase_damping", is_unitary=False)


channels: List[str] = ["amplitudedamping", "depolarizing", "generaldepolarizing", "isotropicdepolarizing", "phasedamping", "reset"]


def kraus_identity_check(kraus: Sequence[Gate]) -> None:  # channels.py:490-518
    d = gates.reshapem(kraus[0].tensor).shape[-1]
    acc = np.zeros((d, d), dtype=np.complex128)
    for k in kraus:
        m = np.asarray(gates.reshapem(k.tensor))
        acc = acc + m.conj().T @ m
    np.testing.assert_allclose(acc, np.eye(d), atol=1e-5)
