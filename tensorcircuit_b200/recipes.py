"""Benchmark / validation circuits of SURVEY.md 8(d) as ``(name, qubits, params)`` gate lists,
plus the TFIM Pauli strings of examples/vqe_parallel_pmap.py:28-34.  Pure host code; the
recipes are seeded so that every run (and the oracle in the tests) builds identical circuits."""

from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np

GateList = List[Tuple[str, Tuple[int, ...], Dict[str, Any]]]


def hea_circuit(n: int, params: np.ndarray) -> GateList:
    """Config 1/3: per layer rx on all, rzz ladder, cnot ladder; params [depth, 2, n]."""
    ops: GateList = []
    for l in range(params.shape[0]):
        for i in range(n):
            ops.append(("rx", (i,), {"theta": float(params[l, 0, i])}))
        for i in range(n - 1):
            ops.append(("rzz", (i, i + 1), {"theta": float(params[l, 1, i])}))
        for i in range(n - 1):
            ops.append(("cnot", (i, i + 1), {}))
    return ops


def tfim_vqe_circuit(n: int, params: np.ndarray) -> GateList:
    """Config 2: H on all, then per layer (rzz ladder; rx all); params [2*layers, n]
    (templates/blocks.py:141-152 shape)."""
    ops: GateList = [("h", (i,), {}) for i in range(n)]
    for l in range(params.shape[0] // 2):
        for i in range(n - 1):
            ops.append(("rzz", (i, i + 1), {"theta": float(params[2 * l, i])}))
        for i in range(n):
            ops.append(("rx", (i,), {"theta": float(params[2 * l + 1, i])}))
    return ops


def random_circuit(n: int, depth: int, seed: int) -> GateList:
    """Config 4/5: each layer = r(theta, alpha, phi) on every qubit (gates.py:545-573), then cnot
    on a random perfect matching (cf. examples/sample_benchmark.py:14-22)."""
    rng = np.random.default_rng(seed)
    ops: GateList = []
    for _ in range(depth):
        ang = rng.uniform(0, 2 * np.pi, size=(n, 3))
        for i in range(n):
            ops.append(("r", (i,), {"theta": float(ang[i, 0]), "alpha": float(ang[i, 1]), "phi": float(ang[i, 2])}))
        perm = rng.permutation(n)
        for j in range(n // 2):
            ops.append(("cnot", (int(perm[2 * j]), int(perm[2 * j + 1])), {}))
    return ops


def tfim_terms(n: int, periodic: bool = True) -> List[Tuple[float, List[int]]]:
    """X_i (weight -1) and Z_i Z_{(i+1)%n} (weight +1) as (weight, ps)."""
    terms = []
    for i in range(n):
        ps = [0] * n
        ps[i] = 1
        terms.append((-1.0, ps))
    for i in range(n if periodic else n - 1):
        ps = [0] * n
        ps[i] = 3
        ps[(i + 1) % n] = 3
        terms.append((1.0, ps))
    return terms


def build(circuit: Any, ops: GateList) -> Any:
    """Replay a gate list through the public gate methods of ``circuit``."""
    for name, q, p in ops:
        getattr(circuit, name)(*q, **p)
    return circuit
