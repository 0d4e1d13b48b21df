"""CPU emulation of the CUDA kernel bodies vs the oracle (no GPU needed).

tests/emu/emu.cu executes the same __host__ __device__ staging / group / expectation code the
sm_100a kernels execute, plus the real host-side planning (make_geom, make_group_map), thread by
thread.  This pins index arithmetic, swizzle and lane planning on the CPU; the -m gpu tests then
check the real kernels through the C ABI."""

import ctypes
import itertools
import os
import subprocess

import numpy as np
import pytest

from oracle import tc_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libtcb200emu.so")


def _build_emu():
    # apply.cu is compiled with -DTCB200_EMU only here: that adds the CPU execution of the
    # register-tile pass (same parameter block and device functions as rpass_kernel)
    srcs = [os.path.join(EMU_DIR, "emu.cu"), os.path.join(ROOT, "tensorcircuit_b200", "csrc", "abi.cu"),
            os.path.join(ROOT, "tensorcircuit_b200", "csrc", "apply.cu"), os.path.join(ROOT, "tensorcircuit_b200", "csrc", "lpass.cu"),
            os.path.join(ROOT, "tensorcircuit_b200", "csrc", "tpass.cu"),
            os.path.join(ROOT, "tensorcircuit_b200", "csrc", "expect.cu")]
    deps = srcs + [os.path.join(ROOT, "tensorcircuit_b200", "csrc", "common.cuh"), os.path.join(ROOT, "tensorcircuit_b200", "csrc", "lpass.cuh"), os.path.join(ROOT, "include", "tcb200.h")]
    if os.path.exists(EMU_LIB) and all(os.path.getmtime(EMU_LIB) > os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    cmd = [nvcc, "-O1", "-std=c++17", "-DTCB200_EMU", "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tensorcircuit_b200", "csrc"), "-gencode", "arch=compute_100a,code=sm_100a"] + srcs + ["-o", EMU_LIB, "-lcudart"]
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


@pytest.fixture(scope="module")
def emu():
    try:
        _build_emu()
    except (subprocess.CalledProcessError, FileNotFoundError) as e:  # pragma: no cover
        pytest.skip("nvcc unavailable for the emulator build: %r" % (e,))
    lib = ctypes.CDLL(EMU_LIB)
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def _ip(a):
    return np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _rand_state(rng, n, dtype):
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


def _bits_to_qubits(n, bits):
    # matrix index bit j <-> bits[j]; oracle wants big-endian qubit order
    return [n - 1 - b for b in reversed(bits)]


def _emu_dense(lib, state, n, bits, mat):
    dt = 0 if state.dtype == np.complex64 else 1
    m = np.ascontiguousarray(mat, dtype=np.complex128)
    rc = lib.emu_apply_dense(state.ctypes.data_as(ctypes.c_void_p), n, dt, len(bits), _ip(bits), _dp(m.view(np.float64)))
    assert rc == 0, lib.emu_last_error()


TOL = {np.complex64: 2e-6, np.complex128: 1e-13}


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("tile_log2", [7, 9, 15])
def test_emu_dense_all_placements(emu, dtype, tile_log2, monkeypatch):
    monkeypatch.setenv("TCB200_TILE_BYTES_LOG2", str(tile_log2))
    rng = np.random.default_rng(tile_log2)
    n = 9 if tile_log2 < 15 else 8
    for k in range(1, 6):
        combos = list(itertools.combinations(range(n), k))
        if len(combos) > 40:
            idx = rng.choice(len(combos), size=40, replace=False)
            combos = [combos[i] for i in idx]
        for bits in combos:
            psi = _rand_state(rng, n, dtype)
            u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
            ref = orc.apply_gate(psi.astype(np.complex128), u, _bits_to_qubits(n, bits), n)
            got = psi.copy()
            _emu_dense(emu, got, n, list(bits), u)
            err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            assert err < 20 * TOL[dtype], (k, bits, err)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_emu_dense_large_tiles(emu, dtype, monkeypatch):
    # production tile size, several tiles, gathered high bits
    monkeypatch.delenv("TCB200_TILE_BYTES_LOG2", raising=False)
    rng = np.random.default_rng(5)
    n = 15
    for bits in [(0,), (14,), (0, 14), (3, 12, 13), (11, 12, 13, 14), (0, 1, 2, 3), (2, 7, 9, 13, 14), (1, 2, 3, 4, 5), (6, 10)]:
        k = len(bits)
        psi = _rand_state(rng, n, dtype)
        u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
        ref = orc.apply_gate(psi.astype(np.complex128), u, _bits_to_qubits(n, bits), n)
        got = psi.copy()
        _emu_dense(emu, got, n, list(bits), u)
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert err < 20 * TOL[dtype], (bits, err)


def test_emu_tiny_states(emu):
    rng = np.random.default_rng(7)
    for dtype in (np.complex64, np.complex128):
        for n in (1, 2, 3, 4):
            for k in range(1, min(n, 5) + 1):
                for bits in itertools.combinations(range(n), k):
                    psi = _rand_state(rng, n, dtype)
                    u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
                    ref = orc.apply_gate(psi.astype(np.complex128), u, _bits_to_qubits(n, bits), n)
                    got = psi.copy()
                    _emu_dense(emu, got, n, list(bits), u)
                    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 20 * TOL[dtype]


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_emu_conflict_free_lane_plan(emu, dtype):
    """The host lane ordering keeps every shared-memory phase at <= 2-way conflicts for any
    target placement, and conflict-free for the common ones ('swizzle')."""
    dt = 0 if dtype == np.complex64 else 1
    n = 20
    worst = {}
    for k in range(1, 5):
        for bits in itertools.combinations(range(12), k):
            d = emu.emu_conflict_degree(n, dt, k, _ip(bits))
            assert d >= 1
            worst[k] = max(worst.get(k, 1), d)
            assert d <= 2, (bits, d)
        # high targets never conflict
        assert emu.emu_conflict_degree(n, dt, k, _ip(list(range(20 - k, 20)))) == 1
    assert worst[1] == 1 and worst[2] == 1


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("tile_log2", [9, 16])
def test_emu_pass_multi_block(emu, dtype, tile_log2, monkeypatch):
    monkeypatch.setenv("TCB200_PASS_TILE_BYTES_LOG2", str(tile_log2))
    dt = 0 if dtype == np.complex64 else 1
    rng = np.random.default_rng(11)
    T = emu.emu_pass_tile_bits(dt)
    n = T + 3 if tile_log2 == 9 else 14
    for trial in range(6):
        n_hi = int(rng.integers(0, 4)) if n > T else 0
        lrow = T - n_hi if n > T else n
        hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist()) if n_hi else []
        avail = list(range(lrow)) + hi
        nops = int(rng.integers(1, 6))
        ref = _rand_state(rng, n, dtype)
        got = ref.copy()
        ref = ref.astype(np.complex128)
        ks, bl, ml = [], [], []
        for _ in range(nops):
            k = int(rng.integers(1, min(4, len(avail)) + 1))
            bits = sorted(rng.choice(avail, size=k, replace=False).tolist())
            u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
            u /= np.linalg.norm(u, 2)
            ref = orc.apply_gate(ref, u, _bits_to_qubits(n, bits), n)
            ks.append(k)
            bl += bits
            ml.append(np.ascontiguousarray(u, dtype=np.complex128).reshape(-1))
        mats = np.concatenate(ml)
        rc = emu.emu_apply_pass(got.ctypes.data_as(ctypes.c_void_p), n, dt, nops, _ip(ks), _ip(bl), _dp(mats.view(np.float64)), n_hi, _ip(hi if hi else [0]))
        assert rc == 0, emu.emu_last_error()
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert err < 50 * TOL[dtype], (trial, err)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("tile_log2", [8, 15])
def test_emu_expectation(emu, dtype, tile_log2, monkeypatch):
    monkeypatch.setenv("TCB200_EXPECT_TILE_BYTES_LOG2", str(tile_log2))
    dt = 0 if dtype == np.complex64 else 1
    rng = np.random.default_rng(13)
    T = emu.emu_expect_tile_bits(dt)
    n = T + 4 if tile_log2 == 8 else 13
    psi = _rand_state(rng, n, dtype)
    for trial in range(8):
        n_hi = int(rng.integers(0, 4)) if n > T else 0
        lrow = T - n_hi if n > T else n
        hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist()) if n_hi else []
        inside = list(range(lrow)) + hi
        nterms = int(rng.integers(1, 9))
        flips, signs, nys, want = [], [], [], []
        for _ in range(nterms):
            ps = [0] * n
            for q in range(n):
                bit = n - 1 - q
                r = rng.random()
                if bit in inside:
                    ps[q] = int(rng.integers(0, 4)) if r < 0.5 else 0
                else:
                    ps[q] = 3 if r < 0.3 else 0  # outside the tile only I / Z are allowed
            x, y, z = orc.resolve_ps(n, ps=ps)
            f, s, ny = orc.pauli_masks(n, x, y, z)
            flips.append(f)
            signs.append(s)
            nys.append(ny)
            want.append(orc.pauli_expectation(psi.astype(np.complex128), n, x, y, z))
        out = np.zeros(2 * nterms)
        fa = np.array(flips, dtype=np.uint64)
        sa = np.array(signs, dtype=np.uint64)
        rc = emu.emu_expect(psi.ctypes.data_as(ctypes.c_void_p), n, dt, nterms,
                            fa.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), sa.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                            _ip(nys), n_hi, _ip(hi if hi else [0]), _dp(out))
        assert rc == 0, emu.emu_last_error()
        got = out[0::2] + 1j * out[1::2]
        np.testing.assert_allclose(got, np.array(want), atol=50 * TOL[dtype])


# ---- persistent TMA pass (tpass.cu): box decomposition, SWIZZLE_128B layout, 512-thread group walk
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_emu_tpass(emu, dtype):
    dt = 0 if dtype == np.complex64 else 1
    rng = np.random.default_rng(17)
    T = emu.emu_pass_tile_bits(dt)
    emu.emu_apply_tpass.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double), ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_int), ctypes.c_int64]
    for trial in range(10):
        n = int(rng.integers(T + 1, T + 5))
        batch = 1 if trial % 3 else 3
        n_hi = min(int(rng.integers(0, 9)), n - T) if trial else 0   # trial 0: h = 0 (row longer than a box)
        lrow = T - n_hi
        hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist()) if n_hi else []
        avail = list(range(lrow)) + hi
        nops = int(rng.integers(1, 6))
        refs = [_rand_state(rng, n, dtype) for _ in range(batch)]
        got = np.stack(refs).copy()
        refs = [r.astype(np.complex128) for r in refs]
        ks, bl, ml = [], [], []
        for _ in range(nops):
            k = int(rng.integers(1, 5))
            bits = sorted(rng.choice(avail, size=k, replace=False).tolist())
            u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
            u /= np.linalg.norm(u, 2)
            refs = [orc.apply_gate(r, u, _bits_to_qubits(n, bits), n) for r in refs]
            ks.append(k)
            bl += bits
            ml.append(np.ascontiguousarray(u, dtype=np.complex128).reshape(-1))
        mats = np.concatenate(ml)
        rc = emu.emu_apply_tpass(got.ctypes.data_as(ctypes.c_void_p), n, dt, nops, _ip(ks), _ip(bl), _dp(mats.view(np.float64)),
                                 n_hi, _ip(hi if hi else [0]), batch)
        assert rc == 0, (rc, emu.emu_last_error())
        for bi in range(batch):
            err = np.linalg.norm(got[bi] - refs[bi]) / np.linalg.norm(refs[bi])
            assert err < 50 * TOL[dtype], (trial, bi, err)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_emu_tpass_conflicts(emu, dtype):
    """Under the hardware SWIZZLE_128B layout each bank-select bit has two source bits (three in the
    software swizzle), so a block can lose a class: never worse than 2-way, and conflict-free for
    1-bit blocks and for the 2-bit blocks of neighbouring qubits the circuits are made of."""
    dt = 0 if dtype == np.complex64 else 1
    n = 20
    T = emu.emu_pass_tile_bits(dt)
    hi = [15, 17]
    local = list(range(T - 2)) + hi
    hist = {}
    for k in (1, 2, 3):
        for bits in itertools.combinations(local, k):
            d = emu.emu_conflict_degree_tpass(n, dt, k, _ip(bits), 2, _ip(hi))
            assert 1 <= d <= 2, (bits, d)
            hist.setdefault(k, []).append(d)
            if k == 1 or (k == 2 and bits[1] == bits[0] + 1):
                assert d == 1, (bits, d)
    assert np.mean(np.array(hist[2]) == 1) > 0.9


def _random_regtile_pass(rng, n, avail, max_tiles=8):
    """random register tiles (<= 4 bits) of 1-/2-bit gates inside `avail`; returns the ABI arrays
    and the gate list [(bits, u)] in execution order"""
    rt_k, rt_bits, rt_nsub, sub_k, sub_bits, mats, gates = [], [], [], [], [], [], []
    for _ in range(int(rng.integers(1, max_tiles + 1))):
        kt = int(rng.integers(1, 5))
        tb = sorted(rng.choice(avail, size=kt, replace=False).tolist())
        ns = int(rng.integers(1, 5))
        rt_k.append(kt)
        rt_bits += tb
        rt_nsub.append(ns)
        for _ in range(ns):
            k = int(rng.integers(1, min(2, kt) + 1))
            gb = sorted(rng.choice(tb, size=k, replace=False).tolist())
            u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
            u /= np.linalg.norm(u, 2)
            sub_k.append(k)
            sub_bits += gb
            mats.append(np.ascontiguousarray(u, dtype=np.complex128).reshape(-1))
            gates.append((gb, u))
    return rt_k, rt_bits, rt_nsub, sub_k, sub_bits, np.concatenate(mats), gates


def test_emu_trpass(emu):
    """trpass_kernel's compute half: 4-bit register tiles with filler positions, group-offset table,
    gate positions, vec0 / non-vec0 access, on the SWIZZLE_128B box layout."""
    rng = np.random.default_rng(23)
    T = emu.emu_pass_tile_bits(0)
    emu.emu_apply_trpass.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 5 + [
        ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int64]
    for trial in range(12):
        n = int(rng.integers(T + 1, T + 4))
        batch = 1 if trial % 4 else 2
        n_hi = min(int(rng.integers(0, 9)), n - T) if trial else 0
        lrow = T - n_hi
        hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist()) if n_hi else []
        avail = list(range(lrow)) + hi
        if trial == 1:
            avail = avail[1:]  # bit 0 never a target: every tile gets it as a filler (vec0)
        rt_k, rt_bits, rt_nsub, sub_k, sub_bits, mats, gates = _random_regtile_pass(rng, n, avail)
        refs = [_rand_state(rng, n, np.complex64) for _ in range(batch)]
        got = np.stack(refs).copy()
        refs = [r.astype(np.complex128) for r in refs]
        for gb, u in gates:
            refs = [orc.apply_gate(r, u, _bits_to_qubits(n, gb), n) for r in refs]
        rc = emu.emu_apply_trpass(got.ctypes.data_as(ctypes.c_void_p), n, len(rt_k), _ip(rt_k), _ip(rt_bits), _ip(rt_nsub), _ip(sub_k),
                                  _ip(sub_bits), _dp(mats.view(np.float64)), n_hi, _ip(hi if hi else [0]), batch)
        assert rc == 0, (rc, emu.emu_last_error())
        for bi in range(batch):
            err = np.linalg.norm(got[bi] - refs[bi]) / np.linalg.norm(refs[bi])
            assert err < 100 * TOL[np.complex64], (trial, bi, err)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_emu_expect_z(emu, dtype):
    """zexpect_kernel's sign split (chunk / thread / iteration parts of popc(e & mask)) and its
    constant +-1 table, thread by thread on the CPU, against the oracle's closed form."""
    dt = 0 if dtype == np.complex64 else 1
    rng = np.random.default_rng(31)
    n = 15
    psi = _rand_state(rng, n, dtype)
    zt = 32 if dt == 0 else 16
    for nterms in (1, 5, zt):
        masks, want = [], []
        for t in range(nterms):
            m = int(rng.integers(1, 2**n)) if t else (1 << (n - 1)) | 1  # top and bottom bit
            z = [n - 1 - b for b in range(n) if (m >> b) & 1]
            masks.append(m)
            want.append(orc.pauli_expectation(psi.astype(np.complex128), n, [], [], z).real)
        out = np.zeros(nterms)
        ma = np.array(masks, dtype=np.uint64)
        rc = emu.emu_expect_z(psi.ctypes.data_as(ctypes.c_void_p), n, dt, nterms, ma.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), _dp(out))
        assert rc == 0
        np.testing.assert_allclose(out, np.array(want), atol=50 * TOL[dtype])
