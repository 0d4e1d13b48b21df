"""Structure-aware fusion (fusion.fuse_structured) is an exact regrouping: the ordered product of
the fused blocks equals the gate-by-gate product for random circuits of mixed gate classes, the
headline pattern is grouped optimally, and the classes are what the gate pass expects."""

import numpy as np
import pytest

from oracle import tc_oracle as orc
from tensorcircuit_b200 import fusion
from tensorcircuit_b200.fusion import GateOp


def _embed(n, qubits, m):
    """full 2^n x 2^n matrix of gate m on `qubits` (TC big-endian convention)"""
    psi = np.eye(2**n, dtype=np.complex128)
    return np.stack([orc.apply_gate(psi[:, j].copy(), m, list(qubits), n) for j in range(2**n)], axis=1)


def _random_ops(rng, n, count):
    ops = []
    for _ in range(count):
        c = int(rng.integers(0, 9))
        if c == 0:
            ops.append(GateOp((int(rng.integers(n)),), orc.m_r(*rng.uniform(0, 6.28, size=3)), "r"))
        elif c == 1:
            ops.append(GateOp((int(rng.integers(n)),), orc.gate_matrix(["h", "x", "y", "z", "s", "t"][int(rng.integers(6))]), "c1"))
        elif c == 2:
            ops.append(GateOp((int(rng.integers(n)),), orc.m_rz(rng.uniform(0, 6.28)), "rz"))
        elif c == 3 and n >= 3:
            q = tuple(int(x) for x in rng.choice(n, size=3, replace=False))
            ops.append(GateOp(q, orc.gate_matrix("toffoli"), "toffoli"))
        else:
            q = tuple(int(x) for x in rng.choice(n, size=2, replace=False))
            name = ["cnot", "cz", "swap", "iswap", "cy"][int(rng.integers(5))] if c < 7 else None
            if name:
                ops.append(GateOp(q, orc.gate_matrix(name), name))
            elif c == 7:
                ops.append(GateOp(q, orc.gate_matrix("rzz", theta=rng.uniform(0, 6.28)), "rzz"))
            else:
                ops.append(GateOp(q, orc.gate_matrix("rxx", theta=rng.uniform(0, 6.28)), "rxx"))
    return ops


@pytest.mark.parametrize("kmax", [2, 3])
@pytest.mark.parametrize("n", [3, 5, 6])
def test_fused_product_equals_gate_product(n, kmax):
    rng = np.random.default_rng(10 * n + kmax)
    for trial in range(12):
        ops = _random_ops(rng, n, 40)
        want = np.eye(2**n, dtype=np.complex128)
        for op in ops:
            want = _embed(n, op.qubits, np.asarray(op.matrix)) @ want
        blocks = fusion.fuse_structured(ops, n, kmax)
        assert sum(b.ngates for b in blocks) == len(ops)
        got = np.eye(2**n, dtype=np.complex128)
        for b in blocks:
            # block matrix: index bit j <-> bits[j]  ==  big-endian over ascending qubits
            got = _embed(n, b.qubits, b.matrix) @ got
            # ('half' is a planner-only class of 1-qubit gates; blocks carry 'dense' for it)
            assert b.kind == fusion.matrix_kind(b.matrix).replace("half", "dense")
            assert len(b.qubits) <= max(kmax, 3)
        assert np.max(np.abs(got - want)) < 1e-12, trial


def test_headline_pattern_is_grouped_into_double_layers():
    """r on all + cnot matching, depth 20: one 4x4 per pair and two layers, bare cnots in between"""
    from tensorcircuit_b200 import recipes

    n = 34
    ops = [GateOp(q, np.asarray(orc.gate_matrix(name, **p)), name) for name, q, p in recipes.random_circuit(n, 20, 3)]
    blocks = fusion.fuse_structured(ops, n, 2)
    dense = [b for b in blocks if b.kind == "dense"]
    perm = [b for b in blocks if b.kind == "perm"]
    assert len(dense) + len(perm) == len(blocks)
    assert all(len(b.bits) == 2 for b in dense)
    assert len(dense) == 170  # 10 double layers x 17 pairs: 2720 real FMA per amplitude instead of 5440
    assert sum(b.ngates for b in blocks) == len(ops)


def test_kinds():
    assert fusion.matrix_kind(orc.gate_matrix("cnot")) == "perm"
    assert fusion.matrix_kind(orc.gate_matrix("x")) == "perm"
    assert fusion.matrix_kind(orc.gate_matrix("swap")) == "perm"
    assert fusion.matrix_kind(orc.gate_matrix("y")) == "mono"
    assert fusion.matrix_kind(orc.gate_matrix("iswap")) == "mono"
    assert fusion.matrix_kind(orc.gate_matrix("cz")) == "diag"
    assert fusion.matrix_kind(orc.gate_matrix("rzz", theta=0.3)) == "diag"
    assert fusion.matrix_kind(orc.m_rz(0.3)) == "diag"
    # 1-qubit gates whose elements are each real or imaginary: half-cost class of the gate pass
    assert fusion.matrix_kind(orc.gate_matrix("h")) == "half"
    assert fusion.matrix_kind(orc.gate_matrix("rx", theta=0.3)) == "half"
    assert fusion.matrix_kind(orc.gate_matrix("ry", theta=0.3)) == "half"
    assert fusion.matrix_kind(orc.m_r(0.3, 0.4, 0.5)) == "dense"
    assert fusion.matrix_kind(np.stack([orc.gate_matrix("rx", theta=t) for t in (0.1, 0.2, 0.0)])) == "half"
    assert fusion.matrix_kind(np.stack([orc.gate_matrix("rx", theta=0.1), orc.m_r(0.3, 0.4, 0.5)])) == "dense"
    assert fusion.matrix_kind(orc.gate_matrix("toffoli")) == "dense"  # 8x8 permutations are not tracked on the host


@pytest.mark.parametrize("n", [5, 7])
def test_pass_plan_reorders_only_commuting_diagonals(n):
    """plan_passes(block_diag=...) orders a diagonal block only against the non-diagonal blocks around it (an
    rzz ladder is not a dependency chain): the product of the blocks taken in PASS order is still the circuit"""
    rng = np.random.default_rng(100 + n)
    for trial in range(8):
        ops = []
        for _ in range(45):
            c = int(rng.integers(0, 6))
            if c <= 2:  # plenty of diagonal gates sharing qubits
                q = tuple(int(x) for x in rng.choice(n, size=2, replace=False))
                ops.append(GateOp(q, orc.gate_matrix(["rzz", "cz"][c % 2], **({"theta": rng.uniform(0, 6.28)} if c % 2 == 0 else {})), "d2"))
            elif c == 3:
                ops.append(GateOp((int(rng.integers(n)),), orc.m_rz(rng.uniform(0, 6.28)), "rz"))
            elif c == 4:
                ops.append(GateOp((int(rng.integers(n)),), orc.gate_matrix("rx", theta=rng.uniform(0, 6.28)), "rx"))
            else:
                q = tuple(int(x) for x in rng.choice(n, size=2, replace=False))
                ops.append(GateOp(q, orc.gate_matrix("cnot"), "cnot"))
        want = np.eye(2**n, dtype=np.complex128)
        for op in ops:
            want = _embed(n, op.qubits, np.asarray(op.matrix)) @ want
        blocks = fusion.fuse_structured(ops, n, 2)
        diag = [b.kind == "diag" for b in blocks]
        assert any(diag)
        # a 4-bit tile with at most 2 gathered bits forces many passes and a lot of reordering
        passes = fusion.plan_passes([b.bits for b in blocks], n, 4, max_hi=2, max_ops=6, max_pass_k=3, block_diag=diag)
        order = [i for p in passes for i in p.block_ids]
        assert sorted(order) == list(range(len(blocks)))
        got = np.eye(2**n, dtype=np.complex128)
        for i in order:
            got = _embed(n, blocks[i].qubits, blocks[i].matrix) @ got
        assert np.max(np.abs(got - want)) < 1e-12, trial
        plain = fusion.plan_passes([b.bits for b in blocks], n, 4, max_hi=2, max_ops=6, max_pass_k=3)
        assert len(passes) <= len(plain)
