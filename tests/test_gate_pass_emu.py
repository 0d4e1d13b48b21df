"""Structure-aware gate pass (csrc/lpass.cu) on the CPU emulation vs the oracle.

The emulator runs the real host-side classification / round scheduling / index-map tracking
and the same __host__ __device__ round bodies and write-back as lpass_kernel.  Covered: every
micro-op class (2x2, 4x4 on every position pair, 8x8, diagonal tables, merged diagonals),
affine permutations (x, cnot, swap, cnot chains, monomial gates with phases: y, cy, iswap),
non-affine permutations (toffoli -> dense), gathered high bits, tiny tiles, both dtypes."""

import ctypes

import numpy as np
import pytest

from oracle import tc_oracle as orc

from .test_emu_kernels import _bits_to_qubits, _dp, _ip, _rand_state, emu  # noqa: F401

TOL = {np.complex64: 3e-6, np.complex128: 1e-13}


def _run_pass(lib, state, n, gates, tile_hi=()):
    """gates: list of (bits ascending, matrix with index bit j <-> bits[j])"""
    dt = 0 if state.dtype == np.complex64 else 1
    ks = [len(b) for b, _ in gates]
    bits = [x for b, _ in gates for x in b]
    mats = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.complex128).reshape(-1) for _, m in gates]))
    hi = list(tile_hi) if len(tile_hi) else [0]
    info = np.zeros(8, dtype=np.float64)
    rc = lib.emu_apply_gate_pass(state.ctypes.data_as(ctypes.c_void_p), n, dt, len(gates), _ip(ks), _ip(bits), _dp(mats.view(np.float64)),
                                 len(tile_hi), _ip(hi), _dp(info))
    assert rc == 0, lib.emu_last_error()
    return dict(rounds=int(info[0]), nlin=int(info[1]), ndiag=int(info[2]), ndense=int(info[3]), conflicts=int(info[4]),
                vec_rounds=int(info[5]), fma=float(info[6]), mat_elems=int(info[7]))


def _oracle(state, n, gates):
    ref = state.astype(np.complex128)
    for bits, m in gates:
        ref = orc.apply_gate(ref, np.asarray(m, dtype=np.complex128), _bits_to_qubits(n, list(bits)), n)
    return ref


def _bit_matrix(name, **kw):
    """matrix of a named gate with index bit j <-> bits[j] (the oracle's matrices are big-endian in
    the qubit list; a k-qubit gate on ascending bits therefore has its qubit list reversed, which
    _bits_to_qubits undoes -- so the oracle matrix is used as is)"""
    return orc.gate_matrix(name, **kw)


def _rand_u(rng, k):
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, _ = np.linalg.qr(a)
    return q


@pytest.fixture(autouse=True)
def _tile(monkeypatch):
    monkeypatch.delenv("TCB200_PASS_TILE_BYTES_LOG2", raising=False)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_dense_micro_ops_every_position(emu, dtype):
    rng = np.random.default_rng(1)
    n = 10
    for trial in range(12):
        gates = []
        for _ in range(6):
            k = int(rng.integers(1, 4))
            bits = sorted(rng.choice(n, size=k, replace=False).tolist())
            gates.append((bits, _rand_u(rng, k)))
        st = _rand_state(rng, n, dtype)
        ref = _oracle(st, n, gates)
        info = _run_pass(emu, st, n, gates)
        assert info["ndense"] == 6 and info["nlin"] == 0
        assert np.max(np.abs(st - ref)) < TOL[dtype] * 4, trial


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_half_cost_one_bit_gates(emu, dtype):
    """rx / ry / h (every element purely real or purely imaginary) run as LOP_RA / LOP_RB: 4 FMA per amplitude
    instead of 8, on every register position, same numbers as the oracle"""
    rng = np.random.default_rng(21)
    n = 10
    for trial in range(10):
        gates = []
        for _ in range(8):
            name = ["rx", "ry", "h"][int(rng.integers(3))]
            m = _bit_matrix("h") if name == "h" else orc.gate_matrix(name, theta=rng.uniform(0, 6.28))
            gates.append(([int(rng.integers(n))], m))
        st = _rand_state(rng, n, dtype)
        ref = _oracle(st, n, gates)
        info = _run_pass(emu, st, n, gates)
        assert info["ndense"] == 8 and info["fma"] == 8 * 4, info
        assert np.max(np.abs(st - ref)) < TOL[dtype] * 4, trial
    # a general 1-bit gate next to them keeps the full cost
    gates = [([3], orc.gate_matrix("rx", theta=0.4)), ([5], orc.m_r(0.3, 0.4, 0.5))]
    st = _rand_state(rng, n, dtype)
    ref = _oracle(st, n, gates)
    info = _run_pass(emu, st, n, gates)
    assert info["fma"] == 4 + 8
    assert np.max(np.abs(st - ref)) < TOL[dtype] * 4


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_affine_permutations_cost_no_round(emu, dtype):
    rng = np.random.default_rng(2)
    n = 9
    X = _bit_matrix("x")
    CNOT = _bit_matrix("cnot")
    SWAP = _bit_matrix("swap")
    gates = []
    for _ in range(40):
        c = rng.integers(0, 3)
        if c == 0:
            gates.append(([int(rng.integers(n))], X))
        else:
            bits = sorted(rng.choice(n, size=2, replace=False).tolist())
            m = CNOT if c == 1 else SWAP
            if rng.integers(2):  # control on the other bit: conjugate by swap
                m = SWAP @ m @ SWAP
            gates.append((bits, m))
    st = _rand_state(rng, n, dtype)
    ref = _oracle(st, n, gates)
    info = _run_pass(emu, st, n, gates)
    assert info["rounds"] == 0 and info["nlin"] == 40 and info["fma"] == 0
    # a pure permutation: bit-exact
    assert np.array_equal(st, ref.astype(dtype))


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_mixed_structure_vs_oracle(emu, dtype):
    rng = np.random.default_rng(3)
    n = 11
    names1 = ["h", "x", "y", "z", "s", "t", "sd", "td"]
    for trial in range(10):
        gates = []
        for _ in range(30):
            c = int(rng.integers(0, 8))
            if c == 0:
                gates.append(([int(rng.integers(n))], _bit_matrix(names1[int(rng.integers(len(names1)))])))
            elif c == 1:
                gates.append(([int(rng.integers(n))], orc.m_r(*rng.uniform(0, 6.28, size=3))))
            elif c == 2:
                gates.append(([int(rng.integers(n))], orc.m_rz(rng.uniform(0, 6.28))))
            else:
                bits = sorted(rng.choice(n, size=2, replace=False).tolist())
                if c == 3:
                    m = _bit_matrix("cnot")
                elif c == 4:
                    m = orc.gate_matrix("rzz", theta=rng.uniform(0, 6.28))
                elif c == 5:
                    m = _bit_matrix(["cz", "cy", "swap", "iswap"][int(rng.integers(4))])
                elif c == 6:
                    m = _rand_u(rng, 2)
                else:
                    m = orc.gate_matrix("rxx", theta=rng.uniform(0, 6.28))
                gates.append((bits, m))
        st = _rand_state(rng, n, dtype)
        ref = _oracle(st, n, gates)
        _run_pass(emu, st, n, gates)
        assert np.max(np.abs(st - ref)) < TOL[dtype] * 6, trial


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_toffoli_fredkin_and_three_bit_diagonals(emu, dtype):
    rng = np.random.default_rng(4)
    n = 8
    gates = []
    for name in ["toffoli", "fredkin"]:
        bits = sorted(rng.choice(n, size=3, replace=False).tolist())
        gates.append((bits, _bit_matrix(name)))
        gates.append(([int(rng.integers(n))], orc.m_r(*rng.uniform(0, 6.28, size=3))))
    ccz = np.diag([1, 1, 1, 1, 1, 1, 1, -1]).astype(np.complex128)
    gates.append(([1, 4, 6], ccz))
    d4 = np.diag(np.exp(1j * rng.uniform(0, 6.28, size=16)))
    gates.append(([0, 2, 3, 7], d4))
    st = _rand_state(rng, n, dtype)
    ref = _oracle(st, n, gates)
    _run_pass(emu, st, n, gates)
    assert np.max(np.abs(st - ref)) < TOL[dtype] * 4


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("tile_log2", [8, 10, 16])
def test_gathered_bits_and_many_tiles(emu, dtype, tile_log2, monkeypatch):
    monkeypatch.setenv("TCB200_PASS_TILE_BYTES_LOG2", str(tile_log2))
    rng = np.random.default_rng(5 + tile_log2)
    n = 14
    T = emu.emu_pass_tile_bits(0 if dtype == np.complex64 else 1)
    if T >= n:
        tile_hi, inside = [], list(range(n))
    else:
        h = min(3, T - 4)
        tile_hi = sorted(rng.choice(np.arange(T - h, n), size=h, replace=False).tolist())
        inside = list(range(T - h)) + tile_hi
    gates = []
    for _ in range(24):
        c = int(rng.integers(0, 4))
        if c == 0:
            gates.append(([int(rng.choice(inside))], orc.m_r(*rng.uniform(0, 6.28, size=3))))
        else:
            bits = sorted(rng.choice(inside, size=2, replace=False).tolist())
            m = [_bit_matrix("cnot"), _rand_u(rng, 2), orc.gate_matrix("rzz", theta=rng.uniform(0, 6.28))][c - 1]
            gates.append((bits, m))
    st = _rand_state(rng, n, dtype)
    ref = _oracle(st, n, gates)
    _run_pass(emu, st, n, gates, tile_hi)
    assert np.max(np.abs(st - ref)) < TOL[dtype] * 5


def test_headline_pattern_free_cnots(emu):
    """r on both qubits, cnot, r on both: the host fuses these into one 4x4 per pair; the bare
    cnot layers in between are absorbed by the index map -- half the FMAs of gate-by-gate."""
    rng = np.random.default_rng(7)
    n = 12
    gates = []
    perm = rng.permutation(n)
    for j in range(n // 2):
        a, b = sorted((int(perm[2 * j]), int(perm[2 * j + 1])))
        gates.append(([a, b], _rand_u(rng, 2)))
    perm = rng.permutation(n)
    for j in range(n // 2):
        a, b = sorted((int(perm[2 * j]), int(perm[2 * j + 1])))
        gates.append(([a, b], _bit_matrix("cnot")))
    perm2 = rng.permutation(n)
    for j in range(n // 2):
        a, b = sorted((int(perm2[2 * j]), int(perm2[2 * j + 1])))
        gates.append(([a, b], _rand_u(rng, 2)))
    st = _rand_state(rng, n, np.complex64)
    ref = _oracle(st, n, gates)
    info = _run_pass(emu, st, n, gates)
    assert info["nlin"] == n // 2 and info["ndense"] == n
    assert info["rounds"] == n // 2  # two 4x4 blocks per round trip
    assert info["fma"] == 16 * n
    assert np.max(np.abs(st - ref)) < 1e-5


def test_capacity_error_is_reported(emu):
    rng = np.random.default_rng(8)
    n = 10
    gates = [([0, 1], _rand_u(rng, 2)) if i % 2 == 0 else ([1, 2], _rand_u(rng, 2)) for i in range(200)]
    st = _rand_state(rng, n, np.complex64)
    ks = [2] * len(gates)
    bits = [x for b, _ in gates for x in b]
    mats = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.complex128).reshape(-1) for _, m in gates]))
    rc = emu.emu_apply_gate_pass(st.ctypes.data_as(ctypes.c_void_p), n, 0, len(gates), _ip(ks), _ip(bits), _dp(mats.view(np.float64)), 0, _ip([0]), None)
    assert rc == -4 and b"gate pass" in emu.emu_last_error()


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("n", [7, 14])
def test_batched_gate_pass_vs_oracle(emu, dtype, n):
    """vmap: per-element matrices for rx / rzz (dense and diagonal classes), shared cnot / h; one
    schedule for the whole batch, every element checked against the oracle"""
    rng = np.random.default_rng(20 + n)
    B = 3
    T = emu.emu_pass_tile_bits(0 if dtype == np.complex64 else 1)
    inside = list(range(n)) if n <= T else list(range(T - 2)) + [n - 2, n - 1]
    tile_hi = [] if n <= T else [n - 2, n - 1]
    gates = []  # (bits, [B matrices] or single matrix)
    for _ in range(20):
        c = int(rng.integers(0, 5))
        if c == 0:
            gates.append(([int(rng.choice(inside))], [orc.m_rx(t) for t in rng.uniform(0, 6.28, size=B)]))
        elif c == 1:
            gates.append((sorted(rng.choice(inside, size=2, replace=False).tolist()), [orc.gate_matrix("rzz", theta=t) for t in rng.uniform(0, 6.28, size=B)]))
        elif c == 2:
            gates.append((sorted(rng.choice(inside, size=2, replace=False).tolist()), _bit_matrix("cnot")))
        elif c == 3:
            gates.append(([int(rng.choice(inside))], _bit_matrix("h")))
        else:
            gates.append((sorted(rng.choice(inside, size=2, replace=False).tolist()), [_rand_u(rng, 2) for _ in range(B)]))
    st = np.stack([_rand_state(rng, n, dtype) for _ in range(B)])
    refs = []
    for b in range(B):
        per = [(bits, m[b] if isinstance(m, list) else m) for bits, m in gates]
        refs.append(_oracle(st[b], n, per))
    ks = [len(b) for b, _ in gates]
    bits = [x for b, _ in gates for x in b]
    flags = [1 if isinstance(m, list) else 0 for _, m in gates]
    mats = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.complex128).reshape(-1) for _, m in gates]))
    info = np.zeros(8)
    rc = emu.emu_apply_gate_pass_batched(st.ctypes.data_as(ctypes.c_void_p), n, 0 if dtype == np.complex64 else 1, len(gates), _ip(ks), _ip(bits),
                                         _dp(mats.view(np.float64)), _ip(flags), len(tile_hi), _ip(tile_hi if tile_hi else [0]), B, _dp(info))
    assert rc == 0, emu.emu_last_error()
    assert info[1] >= 1  # the shared cnots were absorbed into the index map
    for b in range(B):
        assert np.max(np.abs(st[b] - refs[b])) < TOL[dtype] * 6, b
