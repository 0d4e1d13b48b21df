"""Parity of the structure-aware gate pass (lpass_kernel, tcb200_apply_gate_pass) on the GPU
against the oracle (``-m gpu``): every micro-op class, affine permutations, gathered bits,
batch > 1, the config-4 recipe through the planner, and full-size properties.

Tolerances (north star): 1e-5 relative for complex64, 1e-11 for complex128."""

import numpy as np
import pytest
import torch

import tensorcircuit_b200 as tc
from oracle import tc_oracle as orc
from tensorcircuit_b200 import _lib, engine, fusion, recipes
from tensorcircuit_b200.engine import DeviceState
from tensorcircuit_b200.fusion import Block, GateOp

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-5, "complex128": 1e-11}


def _rand_state(rng, n):
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return v / np.linalg.norm(v)


def _bits_to_qubits(n, bits):
    return [n - 1 - b for b in reversed(bits)]


def _rand_u(rng, k):
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, _ = np.linalg.qr(a)
    return q


def _block(n, bits, u):
    bits = tuple(bits)
    m = np.asarray(u, dtype=np.complex128)
    return Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=bits, matrix=m, batched=False, ngates=1, kind=fusion.matrix_kind(m))


def _relerr(got, ref):
    return np.linalg.norm(np.asarray(got) - ref) / np.linalg.norm(ref)


def _random_gates(rng, avail, count, classes):
    names1 = ["h", "x", "y", "z", "s", "t", "sd", "td"]
    out = []
    for _ in range(count):
        c = classes[int(rng.integers(len(classes)))]
        if c == "named1":
            out.append(([int(rng.choice(avail))], orc.gate_matrix(names1[int(rng.integers(len(names1)))])))
        elif c == "r":
            out.append(([int(rng.choice(avail))], orc.m_r(*rng.uniform(0, 6.28, size=3))))
        elif c == "rz":
            out.append(([int(rng.choice(avail))], orc.m_rz(rng.uniform(0, 6.28))))
        elif c == "rxy":  # real diagonal + imaginary / real off-diagonal: the half-cost micro-ops
            out.append(([int(rng.choice(avail))], orc.gate_matrix(["rx", "ry"][int(rng.integers(2))], theta=rng.uniform(0, 6.28))))
        elif c == "u3":
            out.append((sorted(rng.choice(avail, size=3, replace=False).tolist()), _rand_u(rng, 3)))
        elif c == "toffoli":
            out.append((sorted(rng.choice(avail, size=3, replace=False).tolist()), orc.gate_matrix("toffoli")))
        else:
            bits = sorted(rng.choice(avail, size=2, replace=False).tolist())
            if c == "cnot":
                m = orc.gate_matrix("cnot")
                if rng.integers(2):
                    sw = orc.gate_matrix("swap")
                    m = sw @ m @ sw
            elif c == "rzz":
                m = orc.gate_matrix("rzz", theta=rng.uniform(0, 6.28))
            elif c == "mono":
                m = orc.gate_matrix(["cz", "cy", "swap", "iswap"][int(rng.integers(4))])
            elif c == "rxx":
                m = orc.gate_matrix("rxx", theta=rng.uniform(0, 6.28))
            else:
                m = _rand_u(rng, 2)
            out.append((bits, m))
    return out


ALL = ["named1", "r", "rz", "rxy", "u3", "toffoli", "cnot", "rzz", "mono", "rxx", "u2"]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [4, 7, 12, 17])
def test_gate_pass_vs_oracle(dtype, n):
    rng = np.random.default_rng(100 + n)
    T = _lib.lib.tcb200_pass_tile_bits(0 if dtype == "complex64" else 1)
    for trial in range(6):
        if n > T:
            n_hi = int(rng.integers(0, 6))
            lrow = T - n_hi
            hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist())
            avail = list(range(lrow)) + hi
        else:
            hi, avail = [], list(range(n))
        classes = [c for c in ALL if not (n < 5 and c in ("u3", "toffoli"))]
        gates = _random_gates(rng, avail, 40, classes)
        psi = _rand_state(rng, n)
        ref = psi.copy()
        for bits, m in gates:
            ref = orc.apply_gate(ref, m, _bits_to_qubits(n, bits), n)
        st = DeviceState(n, dtype)
        st.load(psi)
        before = dict(engine.STATS)
        nl = st.apply_gate_pass([_block(n, b, m) for b, m in gates], hi)
        assert nl >= 1
        assert engine.STATS["gate_pass_rounds"] > before["gate_pass_rounds"]
        err = _relerr(st.buf[0].cpu().numpy(), ref)
        assert err < TOL[dtype] * 3, (trial, err)


def test_gate_pass_permutations_are_bit_exact():
    n = 16
    rng = np.random.default_rng(5)
    gates = _random_gates(rng, list(range(n)), 60, ["cnot"]) + [([int(b)], orc.gate_matrix("x")) for b in rng.choice(n, size=5)]
    psi = _rand_state(rng, n).astype(np.complex64)
    ref = psi.astype(np.complex128)
    for bits, m in gates:
        ref = orc.apply_gate(ref, m, _bits_to_qubits(n, bits), n)
    st = DeviceState(n, "complex64")
    st.load(psi)
    before = engine.STATS["gate_pass_rounds"]
    hi = fusion.tile_hi_fixpoint(list(range(n)), 13, n)
    # all 16 bits cannot sit in one 13-bit tile: run through the planner instead
    st.apply_planned([_block(n, b, m) for b, m in gates])
    assert engine.STATS["gate_pass_rounds"] == before  # no round trip at all: index maps only
    assert np.array_equal(st.buf[0].cpu().numpy(), ref.astype(np.complex64))
    del hi


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_gate_pass_batch_rows(dtype):
    """batch > 1 with shared matrices: every row is transformed by the same pass"""
    n, B = 13, 3
    rng = np.random.default_rng(6)
    gates = _random_gates(rng, list(range(n)), 25, ["r", "cnot", "rzz", "u2"])
    rows = [_rand_state(rng, n) for _ in range(B)]
    st = DeviceState(n, dtype, batch=B)
    cd = torch.complex64 if dtype == "complex64" else torch.complex128
    st.buf.copy_(torch.from_numpy(np.stack(rows)).to(st.device, dtype=cd))
    st.apply_planned([_block(n, b, m) for b, m in gates])
    for r in range(B):
        ref = rows[r]
        for bits, m in gates:
            ref = orc.apply_gate(ref, m, _bits_to_qubits(n, bits), n)
        assert _relerr(st.buf[r].cpu().numpy(), ref) < TOL[dtype] * 3


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_config4_recipe_structured_vs_oracle(dtype):
    """config-4 recipe (r on all + cnot matching) at an oracle-checkable size through
    fuse_structured + the pass planner + the gate pass"""
    n, depth = 20, 8
    rc = recipes.random_circuit(n, depth, 11)
    ops = []
    for name, q, p in rc:
        ops.append(GateOp(q, np.asarray(orc.gate_matrix(name, **p)), name))
    blocks = fusion.fuse_structured(ops, n, 2)
    assert any(b.kind == "perm" for b in blocks)
    st = DeviceState(n, dtype)
    st.init_zero()
    st.apply_planned(blocks)
    ref = orc.run_gatelist(n, rc).state()
    assert _relerr(st.buf[0].cpu().numpy(), ref) < TOL[dtype]


def test_gate_pass_matches_dense_pass_at_n26():
    """size-independent property at a size the oracle cannot reach: the gate pass and the older
    dense multi-block pass (cpass) agree on the config-4 recipe"""
    n, depth = 26, 6
    rc = recipes.random_circuit(n, depth, 5)
    c1 = tc.Circuit(n)
    recipes.build(c1, rc)
    s1 = c1._ensure_state()
    old = DeviceState.use_gate_pass
    try:
        DeviceState.use_gate_pass = False
        c2 = tc.Circuit(n)
        recipes.build(c2, rc)
        s2 = c2._ensure_state()
    finally:
        DeviceState.use_gate_pass = old
    d = (s1.buf - s2.buf).abs().max().item()
    assert d < 2e-6, d
    assert abs(float(s1.norm2()[0]) - 1.0) < 1e-5


@pytest.mark.parametrize("mode", ["1", "2", "tma"])
def test_pipelined_gate_pass_is_bit_identical(mode, monkeypatch):
    """the opt-in staging variants -- persistent pipelines (lpass_pipe_kernel / lpass_pipe2_kernel: three tile
    buffers per SM, mbarrier hand-over) and TMA-staged tiles (lpass_tma_kernel: SWIZZLE_128B layout as the
    start of the index map) -- run the same arithmetic as lpass_fast_kernel: identical bits, vs the oracle too"""
    n, depth = 20, 6
    rc = recipes.random_circuit(n, depth, 7)
    ops = [GateOp(q, np.asarray(orc.gate_matrix(name, **p)), name) for name, q, p in rc]
    blocks = fusion.fuse_structured(ops, n, 2)

    def run():
        st = DeviceState(n, "complex64")
        st.init_zero()
        st.apply_planned(blocks)
        return st.buf.clone()

    monkeypatch.delenv("TCB200_PIPE", raising=False)
    monkeypatch.delenv("TCB200_GATE_TMA", raising=False)
    base = run()
    for grid in ("3", "77", "1000"):  # many tiles per CTA, uneven split, more CTAs than tiles
        if mode == "tma":
            monkeypatch.setenv("TCB200_GATE_TMA", "1")
        else:
            monkeypatch.setenv("TCB200_PIPE", mode)
        monkeypatch.setenv("TCB200_PIPE_MIN_TILES", "1")
        monkeypatch.setenv("TCB200_PIPE_GRID", grid)
        got = run()
        assert torch.equal(got, base), (mode, grid)
    ref = orc.run_gatelist(n, rc).state()
    assert _relerr(base[0].cpu().numpy(), ref) < TOL["complex64"]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [6, 15])
def test_batched_gate_pass_vmap_vs_oracle(dtype, n):
    """backend.vmap over a HEA layer stack: per-element rx / rzz matrices, shared cnots, through
    fuse_structured + planner + tcb200_apply_gate_pass_batched; every element vs the oracle"""
    tc.set_dtype(dtype)
    try:
        B, depth = 5, 3
        rng = np.random.default_rng(n)
        params = rng.uniform(0, 2 * np.pi, size=(B, depth, 2, n))

        def f(p):
            c = tc.Circuit(n)
            for l in range(depth):
                for i in range(n):
                    c.rx(i, theta=p[l, 0, i])
                for i in range(n - 1):
                    c.rzz(i, i + 1, theta=p[l, 1, i])
                for i in range(n - 1):
                    c.cnot(i, i + 1)
            return c.state()

        before = engine.STATS["gate_pass_rounds"]
        got = np.asarray(tc.backend.vmap(f)(params))
        assert engine.STATS["gate_pass_rounds"] > before
        for b in range(B):
            ref = orc.run_gatelist(n, recipes.hea_circuit(n, params[b])).state()
            assert _relerr(got[b], ref) < TOL[dtype] * 2, b
    finally:
        tc.set_dtype("complex64")
