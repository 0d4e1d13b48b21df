"""Parity of the sm_100a kernels, called through the C ABI, against the oracle (``-m gpu``).

Tolerances (north star): amplitudes / expectations within 1e-5 relative for complex64 and
1e-11 for complex128; sample indices identical to the reference rule evaluated in float64
except at CDF ties within tolerance."""

import itertools

import numpy as np
import pytest
import torch

import tensorcircuit_b200 as tc
from oracle import tc_oracle as orc
from tensorcircuit_b200 import _lib, engine
from tensorcircuit_b200.engine import DeviceState
from tensorcircuit_b200.fusion import Block

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-5, "complex128": 1e-11}


def _rand_state(rng, n):
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return v / np.linalg.norm(v)


def _bits_to_qubits(n, bits):
    return [n - 1 - b for b in reversed(bits)]


def _block(n, bits, u):
    bits = tuple(bits)
    return Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=bits, matrix=np.asarray(u, dtype=np.complex128), batched=False, ngates=1)


def _relerr(got, ref):
    return np.linalg.norm(np.asarray(got) - ref) / np.linalg.norm(ref)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [13, 18])
def test_apply_dense_placements(dtype, n):
    rng = np.random.default_rng(n)
    psi = _rand_state(rng, n)
    placements = []
    for k in range(1, 6):
        placements += [tuple(range(k)), tuple(range(n - k, n))]
        for _ in range(6):
            placements.append(tuple(sorted(rng.choice(n, size=k, replace=False).tolist())))
    for bits in placements:
        k = len(bits)
        u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
        u /= np.linalg.norm(u, 2)
        st = DeviceState(n, dtype)
        st.load(psi)
        st.apply_block(_block(n, bits, u))
        ref = orc.apply_gate(psi, u, _bits_to_qubits(n, bits), n)
        err = _relerr(st.buf[0].cpu().numpy(), ref)
        assert err < TOL[dtype], (bits, err)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_apply_dense_every_low_placement(dtype):
    """every k-subset of the 7 lowest bits (+ one high bit): the swizzle / lane-plan cases"""
    n = 14
    rng = np.random.default_rng(1)
    psi = _rand_state(rng, n)
    st = DeviceState(n, dtype)
    for k in (1, 2, 3, 4):
        for low in itertools.combinations(range(7), k - 1 if k > 1 else 1):
            bits = tuple(low) + ((13,) if k > 1 else ())
            kk = len(bits)
            u = rng.normal(size=(2**kk, 2**kk)) + 1j * rng.normal(size=(2**kk, 2**kk))
            u /= np.linalg.norm(u, 2)
            st.load(psi)
            st.apply_block(_block(n, bits, u))
            ref = orc.apply_gate(psi, u, _bits_to_qubits(n, bits), n)
            assert _relerr(st.buf[0].cpu().numpy(), ref) < TOL[dtype], bits


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_tiny_states(dtype):
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 5):
        for k in range(1, min(n, 5) + 1):
            bits = tuple(sorted(rng.choice(n, size=k, replace=False).tolist()))
            psi = _rand_state(rng, n)
            u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
            st = DeviceState(n, dtype)
            st.load(psi)
            st.apply_block(_block(n, bits, u))
            ref = orc.apply_gate(psi, u, _bits_to_qubits(n, bits), n)
            assert _relerr(st.buf[0].cpu().numpy(), ref) < 10 * TOL[dtype], (n, bits)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_apply_pass_multi_block(dtype):
    n = 17
    rng = np.random.default_rng(3)
    T = _lib.lib.tcb200_pass_tile_bits(0 if dtype == "complex64" else 1)
    for trial in range(8):
        n_hi = int(rng.integers(0, 5))
        lrow = T - n_hi
        hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist())
        avail = list(range(lrow)) + hi
        psi = _rand_state(rng, n)
        ref = psi.copy()
        blocks = []
        for _ in range(int(rng.integers(1, 9))):
            k = int(rng.integers(1, 5))
            bits = sorted(rng.choice(avail, size=k, replace=False).tolist())
            u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
            u /= np.linalg.norm(u, 2)
            ref = orc.apply_gate(ref, u, _bits_to_qubits(n, bits), n)
            blocks.append(_block(n, bits, u))
        st = DeviceState(n, dtype)
        st.load(psi)
        st.apply_pass(blocks, hi)
        assert _relerr(st.buf[0].cpu().numpy(), ref) < 5 * TOL[dtype], trial


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_apply_batched(dtype):
    n, B = 12, 5
    rng = np.random.default_rng(4)
    psi = _rand_state(rng, n)
    st = DeviceState(n, dtype, batch=B)
    st.load(psi)
    refs = [psi.copy() for _ in range(B)]
    for bits in [(0, 3), (11,), (2, 5, 9, 10), (1, 4, 6)]:
        k = len(bits)
        us = rng.normal(size=(B, 2**k, 2**k)) + 1j * rng.normal(size=(B, 2**k, 2**k))
        us /= np.linalg.norm(us, 2, axis=(1, 2))[:, None, None]
        blk = _block(n, bits, us)
        blk.batched = True
        st.apply_block(blk)
        refs = [orc.apply_gate(refs[b], us[b], _bits_to_qubits(n, bits), n) for b in range(B)]
    # shared matrix on a batched state
    u = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    u /= np.linalg.norm(u, 2)
    st.apply_block(_block(n, (1, 7), u))
    refs = [orc.apply_gate(r, u, _bits_to_qubits(n, (1, 7)), n) for r in refs]
    got = st.buf.cpu().numpy()
    for b in range(B):
        assert _relerr(got[b], refs[b]) < 5 * TOL[dtype]
    n2 = st.norm2()
    np.testing.assert_allclose(n2, [np.vdot(r, r).real for r in refs], rtol=20 * TOL[dtype])


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_apply_diag(dtype):
    n = 15
    rng = np.random.default_rng(5)
    psi = _rand_state(rng, n)
    dt = 0 if dtype == "complex64" else 1
    for bits in [(0,), (3, 14), (0, 1, 2, 3, 4, 5, 6, 7, 8, 9), (2, 6, 11, 13)]:
        k = len(bits)
        d = np.exp(1j * rng.uniform(0, 6, size=2**k))
        st = DeviceState(n, dtype)
        st.load(psi)
        b = np.asarray(bits, dtype=np.int32)
        _lib.check(_lib.lib.tcb200_apply_diag(engine._ptr(st.buf), n, dt, k, _lib.iptr(b), _lib.dptr(np.ascontiguousarray(d).view(np.float64)), None, 1, engine._stream()))
        ref = orc.apply_gate(psi, np.diag(d), _bits_to_qubits(n, bits), n)
        assert _relerr(st.buf[0].cpu().numpy(), ref) < TOL[dtype], bits


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_expectation_kernel(dtype):
    n = 16
    rng = np.random.default_rng(6)
    psi = _rand_state(rng, n)
    st = DeviceState(n, dtype)
    st.load(psi)
    pss = [list(r) for r in rng.integers(0, 4, size=[5, n])]
    # sparse strings incl. high flips and Y's
    for q in (0, 1, 7, 15):
        for p in (1, 2, 3):
            ps = [0] * n
            ps[q] = p
            pss.append(ps)
    pss.append([3] * n)
    pss.append([0] * n)
    fl, sg, ny, want = [], [], [], []
    for ps in pss[5:]:
        x, y, z = orc.resolve_ps(n, ps=ps)
        f, s, c = orc.pauli_masks(n, x, y, z)
        fl.append(f); sg.append(s); ny.append(c)
        want.append(orc.pauli_expectation(psi, n, x, y, z))
    got = st.expectation_terms(fl, sg, ny)[0]
    np.testing.assert_allclose(got, want, atol=TOL[dtype])
    # dense random strings flip many high bits: must raise cleanly (too many gathered bits) or be right
    c = tc.Circuit(6, inputs=_rand_state(rng, 6))
    tc.set_dtype(dtype)
    c = tc.Circuit(6, inputs=psi[:64] / np.linalg.norm(psi[:64]))
    for ps in rng.integers(0, 4, size=[6, 6]):
        o = orc.OracleCircuit(6, inputs=psi[:64] / np.linalg.norm(psi[:64]))
        np.testing.assert_allclose(c.expectation_ps(ps=list(ps)), o.expectation_ps(ps=list(ps)), atol=TOL[dtype])
    tc.set_dtype("complex64")
    # determinism: run-to-run identical bits
    a = st.expectation_terms(fl, sg, ny)
    b = st.expectation_terms(fl, sg, ny)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [5, 14, 20])
def test_sampler_rule(dtype, n):
    rng = np.random.default_rng(7 + n)
    psi = _rand_state(rng, n)
    if n == 14:  # zero-probability runs and a heavy entry
        psi[: 2**10] = 0
        psi[5000] = 3.0
        psi /= np.linalg.norm(psi)
    st = DeviceState(n, dtype)
    st.load(psi)
    u = rng.random(20000)
    u[:4] = [0.0, 0.5, 1 - 1e-12, 0.25]
    got, total = st.sample(u, return_total=True)
    dev = st.buf[0].cpu().numpy().astype(np.complex128)
    p = np.abs(dev) ** 2
    np.testing.assert_allclose(total, p.sum(), rtol=1e-12)
    cdf = np.cumsum(p)
    r = cdf[-1] * (1 - u)
    want = np.searchsorted(cdf, r, side="left")
    want = np.minimum(want, 2**n - 1)
    bad = np.nonzero(got != want)[0]
    assert len(bad) < 20
    for i in bad:  # ties only
        lo, hi = sorted((int(got[i]), int(want[i])))
        assert abs(cdf[lo] - r[i]) <= 1e-12 * cdf[-1] + 1e-15, (i, got[i], want[i])
    assert np.all(p[got] > 0)
    # chi^2-free sanity: empirical mean of bit 0
    assert abs(np.mean(got & 1) - p[1::2].sum() / p.sum()) < 0.02


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_norm_and_probability(dtype):
    n = 19
    rng = np.random.default_rng(8)
    psi = 3.0 * _rand_state(rng, n)
    st = DeviceState(n, dtype)
    st.load(psi)
    np.testing.assert_allclose(st.norm2()[0], 9.0, rtol=1e-6 if dtype == "complex64" else 1e-13)
    p = st.probability()[0].cpu().numpy()
    np.testing.assert_allclose(p, np.abs(psi) ** 2, rtol=1e-5, atol=1e-12)
    st.init_zero()
    v = st.buf[0].cpu().numpy()
    assert v[0] == 1 and np.count_nonzero(v) == 1


def test_run_circuit_host():
    """Host-buffer entry point: gate blocks + sampling in one C call."""
    n = 12
    rng = np.random.default_rng(9)
    st = DeviceState(n, "complex64")
    ks, bits, mats = [], [], []
    ref = np.zeros(2**n, dtype=np.complex128)
    ref[0] = 1
    for _ in range(10):
        k = int(rng.integers(1, 5))
        b = sorted(rng.choice(n, size=k, replace=False).tolist())
        u = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))[0]
        ref = orc.apply_gate(ref, u, _bits_to_qubits(n, b), n)
        ks.append(k); bits += b; mats.append(np.ascontiguousarray(u, dtype=np.complex128).reshape(-1))
    shots = 1000
    u01 = rng.random(shots)
    out = np.zeros(shots, dtype=np.int64)
    need = _lib.lib.tcb200_sample_workspace_bytes(n) + shots * 16 + 4096
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    ka = np.asarray(ks, dtype=np.int32); ba = np.asarray(bits, dtype=np.int32); ma = np.concatenate(mats)
    _lib.check(_lib.lib.tcb200_run_circuit_host(engine._ptr(st.buf), n, 0, 1, len(ks), _lib.iptr(ka), _lib.iptr(ba), _lib.dptr(ma.view(np.float64)),
                                                shots, _lib.dptr(u01), _lib.i64ptr(out), engine._ptr(ws), ws.numel(), engine._stream()))
    assert _relerr(st.buf[0].cpu().numpy(), ref) < 1e-5
    cdf = orc.sample_cdf(np.abs(st.buf[0].cpu().numpy().astype(np.complex128)) ** 2)
    want = np.searchsorted(cdf, cdf[-1] * (1 - u01), side="left")
    assert np.mean(out == want) > 0.995


def test_config4_recipe_n22_and_properties():
    """Config 4 recipe (random circuit, seed 3) at n = 22 vs the oracle, then size-independent
    properties at n = 28: norm preservation and circuit followed by its inverse = identity."""
    n = 22
    ops = orc.random_circuit(n, 3, seed=3)
    c = tc.Circuit(n)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    o = orc.run_gatelist(n, ops)
    psi = np.asarray(c.state())
    assert _relerr(psi, o.state()) < 1e-5
    zs = c.expectation_ps_many([[3 if j == i else 0 for j in range(n)] for i in range(n)])
    want = [o.expectation_ps(z=[i]).real for i in range(n)]
    np.testing.assert_allclose(np.real(zs), want, atol=1e-5)
    n = 28
    ops = orc.random_circuit(n, 4, seed=5)
    c = tc.Circuit(n)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    st = c._ensure_state()
    np.testing.assert_allclose(st.norm2()[0], 1.0, atol=2e-5)
    # sample marginals agree with <Z_i> from the expectation kernel
    u = np.random.default_rng(4).random(200000)
    s = c.sample(batch=len(u), allow_state=True, format="sample_int", status=u)
    zs = np.real(c.expectation_ps_many([[3 if j == i else 0 for j in range(n)] for i in range(n)]))
    for i in range(n):
        emp = np.mean(1 - 2 * ((s >> (n - 1 - i)) & 1))
        assert abs(emp - zs[i]) < 0.02
    c.append(c.inverse())
    st = c._ensure_state()
    amp0 = c.amplitude("0" * n)
    assert abs(abs(amp0) - 1.0) < 1e-4


def test_config2_tfim_energy_n20():
    """Config 2 recipe at n = 20 vs the oracle (energy of 2n TFIM strings)."""
    n = 20
    params = np.random.default_rng(1).uniform(0, 2 * np.pi, [8, n])
    ops = orc.tfim_vqe_circuit(n, params)
    c = tc.Circuit(n)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    o = orc.run_gatelist(n, ops)
    terms = orc.tfim_terms(n)
    e = tc.templates.measurements.pauli_sum_expectation(c, [ps for _, ps in terms], [w for w, _ in terms])
    want = sum(w * o.expectation_ps(ps=ps).real for w, ps in terms)
    assert abs(e - want) / max(abs(want), 1e-3 * len(terms)) < 1e-5


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("kf,regtiles", [(2, False), (2, True), (3, False)])
def test_apply_planned_vs_oracle(dtype, kf, regtiles, monkeypatch):
    """Pass planner + multi-block pass kernels (plain and register-tile) on a 21-qubit circuit."""
    from tensorcircuit_b200.fusion import fuse

    monkeypatch.setattr(DeviceState, "use_regtiles", regtiles)
    n = 21
    ops = orc.random_circuit(n, 4, seed=11)
    tc.set_dtype(dtype)
    c = tc.Circuit(n)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    blocks = fuse(c._ops, n, kmax=kf)
    st = DeviceState(n, dtype)
    st.init_zero()
    npass = st.apply_planned(blocks)
    assert npass < len(blocks)
    o = orc.run_gatelist(n, ops)
    assert _relerr(st.buf[0].cpu().numpy(), o.state()) < TOL[dtype]
    tc.set_dtype("complex64")


def test_config2_full_size_n28():
    """Config 2 at BASELINE's size (n = 28, 56 TFIM strings): complex64 against the engine's own
    complex128 run (the oracle cannot hold 2^28 here), plus norm and Hermiticity properties."""
    from tensorcircuit_b200 import recipes

    n = 28
    params = np.random.default_rng(1).uniform(0, 2 * np.pi, [8, n])
    ops = recipes.tfim_vqe_circuit(n, params)
    terms = recipes.tfim_terms(n)
    pss = [ps for _, ps in terms]
    vals = {}
    for dtype in ("complex64", "complex128"):
        tc.set_dtype(dtype)
        c = recipes.build(tc.Circuit(n), ops)
        v = np.asarray(c.expectation_ps_many(pss))
        assert abs(c._ensure_state().norm2()[0] - 1.0) < (1e-5 if dtype == "complex64" else 1e-11)
        assert np.max(np.abs(v.imag)) < (1e-5 if dtype == "complex64" else 1e-11)
        vals[dtype] = v.real
        del c
    tc.set_dtype("complex64")
    e64 = sum(w * v for (w, _), v in zip(terms, vals["complex64"]))
    e128 = sum(w * v for (w, _), v in zip(terms, vals["complex128"]))
    assert np.max(np.abs(vals["complex64"] - vals["complex128"])) < 1e-5
    assert abs(e64 - e128) / max(abs(e128), 1e-3 * len(terms)) < 1e-5


def test_config3_full_size_vmap_1024x20():
    """Config 3 at BASELINE's size: vmap over 1024 parameter sets of a 20-qubit HEA (8 GiB batched
    state), TFIM energy; a 12-element subsample is checked against the oracle."""
    from tensorcircuit_b200 import recipes

    n, B, depth = 20, 1024, 4
    params = np.random.default_rng(2).uniform(0, 2 * np.pi, size=[B, depth, 2, n])
    terms = recipes.tfim_terms(n)
    pss = [ps for _, ps in terms]
    ws = [w for w, _ in terms]

    def energy(p):
        c = tc.Circuit(n)
        for l in range(depth):
            for i in range(n):
                c.rx(i, theta=p[l, 0, i])
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=p[l, 1, i])
            for i in range(n - 1):
                c.cnot(i, i + 1)
        return tc.templates.measurements.pauli_sum_expectation(c, pss, ws)

    got = tc.backend.vmap(energy)(params)
    assert got.shape == (B,)
    for b in np.random.default_rng(0).choice(B, size=12, replace=False):
        o = orc.run_gatelist(n, orc.hea_circuit(n, params[b]))
        want = sum(w * o.expectation_ps(ps=ps).real for w, ps in terms)
        assert abs(got[b] - want) / max(abs(want), 1e-3 * len(terms)) < 1e-5, b


# ---- persistent TMA pipeline (tpass.cu) ------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_tma_pass_vs_oracle_and_staged_kernel(dtype, monkeypatch):
    """tcb200_apply_pass_host through tpass_kernel (default) and through cpass_kernel
    (TCB200_TMA=0): both against the oracle, and bit-identical to each other (same FMA order)."""
    monkeypatch.setenv("TCB200_TMA_STRICT", "1")
    rng = np.random.default_rng(5)
    T = _lib.lib.tcb200_pass_tile_bits(0 if dtype == "complex64" else 1)
    for trial in range(10):
        n = int(rng.integers(T + 1, T + 7))
        batch = 1 if trial % 3 else 3
        n_hi = min(int(rng.integers(0, 9)), n - T) if trial else 0
        lrow = T - n_hi
        hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist())
        avail = list(range(lrow)) + hi
        psis = [_rand_state(rng, n) for _ in range(batch)]
        refs = [p.copy() for p in psis]
        blocks = []
        mat_bytes = 0
        for _ in range(int(rng.integers(1, 9))):
            k = int(rng.integers(1, 5))
            mat_bytes += 4**k * (8 if dtype == "complex64" else 16)
            if mat_bytes > 12 * 1024:  # the pass blob of matrices is limited to 12 KiB
                break
            bits = sorted(rng.choice(avail, size=k, replace=False).tolist())
            u = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
            u /= np.linalg.norm(u, 2)
            refs = [orc.apply_gate(r, u, _bits_to_qubits(n, bits), n) for r in refs]
            blocks.append(_block(n, bits, u))
        outs = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("TCB200_TMA", mode)
            st = DeviceState(n, dtype, batch=batch)
            for b in range(batch):
                st.buf[b].copy_(torch.as_tensor(psis[b]).to(st.buf.dtype))
            c0 = _lib.lib.tcb200_tma_pass_count()
            st.apply_pass_host(blocks, hi)
            torch.cuda.synchronize()
            assert _lib.lib.tcb200_tma_pass_count() - c0 == (1 if mode == "1" else 0), (trial, mode)
            outs[mode] = st.buf.cpu().numpy()
            for b in range(batch):
                assert _relerr(outs[mode][b], refs[b]) < 5 * TOL[dtype], (trial, mode, b)
        if dtype == "complex128":
            assert np.array_equal(outs["0"], outs["1"]), trial
        else:  # complex64 passes of narrow blocks run as register tiles: same math, other schedule
            assert np.abs(outs["0"] - outs["1"]).max() < 1e-6, trial


def test_tma_pass_many_tiles_per_cta(monkeypatch):
    """2^24 amplitudes = 2048 tiles over 148 persistent CTAs: the ring wraps ~14 times per CTA
    (mbarrier phases, buffer hand-over between bulk stores and loads)."""
    monkeypatch.setenv("TCB200_TMA", "1")
    monkeypatch.setattr(DeviceState, "use_gate_pass", False)  # the opt-in pipeline sits behind the dense pass
    n, dtype = 24, "complex64"
    c = tc.Circuit(n)
    ops = orc.random_circuit(n, 4, seed=11)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    c0 = _lib.lib.tcb200_tma_pass_count()
    got = np.asarray(c.state())
    assert _lib.lib.tcb200_tma_pass_count() > c0
    o = orc.run_gatelist(n, ops)
    assert _relerr(got, o.state()) < 2e-5


def test_tma_regtile_pass_vs_oracle(monkeypatch):
    """tcb200_apply_rpass_host through trpass_kernel: random <= 4-bit register tiles of 1-/2-bit
    gates (filler positions, vec0 and non-vec0 access, gathered bits, batch > 1)."""
    import ctypes

    monkeypatch.setenv("TCB200_TMA_STRICT", "1")
    monkeypatch.setenv("TCB200_TMA", "1")
    rng = np.random.default_rng(29)
    T = _lib.lib.tcb200_pass_tile_bits(0)
    ip = _lib.iptr
    for trial in range(12):
        n = int(rng.integers(T + 1, T + 7))
        batch = 1 if trial % 4 else 2
        n_hi = min(int(rng.integers(0, 9)), n - T) if trial else 0
        lrow = T - n_hi
        hi = sorted(rng.choice(np.arange(lrow, n), size=n_hi, replace=False).tolist())
        avail = list(range(lrow)) + hi
        if trial == 1:
            avail = avail[1:]
        rt_k, rt_bits, rt_nsub, sub_k, sub_bits, mats, gates = [], [], [], [], [], [], []
        for _ in range(int(rng.integers(1, 9))):
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
        mats = np.concatenate(mats)
        psis = [_rand_state(rng, n) for _ in range(batch)]
        refs = [p.copy() for p in psis]
        for gb, u in gates:
            refs = [orc.apply_gate(r, u, _bits_to_qubits(n, gb), n) for r in refs]
        st = DeviceState(n, "complex64", batch=batch)
        for b in range(batch):
            st.buf[b].copy_(torch.as_tensor(psis[b]).to(st.buf.dtype))
        a = [np.asarray(x, dtype=np.int32) for x in (rt_k, rt_bits, rt_nsub, sub_k, sub_bits, hi if hi else [0])]
        c0 = _lib.lib.tcb200_tma_pass_count()
        _lib.check(_lib.lib.tcb200_apply_rpass_host(ctypes.c_void_p(st.buf.data_ptr()), n, 0, len(rt_k), ip(a[0]), ip(a[1]), ip(a[2]), ip(a[3]),
                                                    ip(a[4]), _lib.dptr(mats.view(np.float64)), n_hi, ip(a[5]), batch, None))
        torch.cuda.synchronize()
        assert _lib.lib.tcb200_tma_pass_count() == c0 + 1
        out = st.buf.cpu().numpy()
        for b in range(batch):
            assert _relerr(out[b], refs[b]) < 10 * TOL["complex64"], (trial, b)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n,batch", [(13, 1), (17, 1), (22, 1), (15, 3)])
def test_expect_z_kernel(dtype, n, batch):
    """Diagonal strings through tcb200_expect_z (one streaming read for up to 32 / 16 strings)
    against the oracle and against the general tile kernel; more strings than one launch holds."""
    rng = np.random.default_rng(n)
    psis = [_rand_state(rng, n) for _ in range(batch)]
    st = DeviceState(n, dtype, batch=batch)
    for b in range(batch):
        st.buf[b].copy_(torch.as_tensor(psis[b]).to(st.buf.dtype))
    nterms = 45
    masks = [int(m) for m in rng.integers(1, 2**n, size=nterms)]
    masks[0] = (1 << (n - 1)) | 1
    masks[1] = (1 << n) - 1
    want = np.zeros((batch, nterms))
    for b in range(batch):
        p = np.abs(psis[b]) ** 2
        idx = np.arange(2**n, dtype=np.uint64)
        for t, m in enumerate(masks):
            par = np.zeros(2**n, dtype=np.uint64)
            x = idx & np.uint64(m)
            for s in range(n):
                par ^= (x >> np.uint64(s)) & np.uint64(1)
            want[b, t] = np.sum(p * (1.0 - 2.0 * par.astype(np.float64)))
    engine.reset_stats()
    got = st.expectation_terms([0] * nterms, masks, [0] * nterms)
    zt = _lib.lib.tcb200_expect_z_max_terms(0 if dtype == "complex64" else 1)
    assert engine.STATS["expect_launches"] == -(-nterms // zt)
    assert np.abs(got.imag).max() == 0.0
    np.testing.assert_allclose(got.real, want, atol=4 * TOL[dtype])
    # mixed with flipping strings: ids are scattered back to their places
    fl = [0, 1, 0, 1 << (n - 1), 0]
    sg = [masks[0], 0, masks[2], 0, masks[3]]
    mix = st.expectation_terms(fl, sg, [0] * 5)
    np.testing.assert_allclose(mix[:, [0, 2, 4]].real, want[:, [0, 2, 3]], atol=4 * TOL[dtype])
    for b in range(batch):
        x0 = 2 * np.real(np.vdot(psis[b][0::2], psis[b][1::2]))
        np.testing.assert_allclose(mix[b, 1].real, x0, atol=4 * TOL[dtype])
    assert np.array_equal(st.expectation_terms([0] * nterms, masks, [0] * nterms), got)  # deterministic


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [1, 4, 13, 21])
def test_masked_norm2(dtype, n):
    """tcb200_masked_norm2: probability mass of a partial measurement record (basecircuit.py:359-443)."""
    rng = np.random.default_rng(40 + n)
    psi = _rand_state(rng, n)
    st = DeviceState(n, dtype)
    st.load(psi)
    p = np.abs(psi) ** 2
    idx = np.arange(2**n, dtype=np.uint64)
    cases = [(0, 0), (1, 0), (1, 1), (1 << (n - 1), 1 << (n - 1)), ((1 << n) - 1, int(rng.integers(0, 2**n)))]
    for _ in range(4):
        m = int(rng.integers(0, 2**n))
        cases.append((m, int(rng.integers(0, 2**n)) & m))
    for mask, value in cases:
        want = float(np.sum(p[(idx & np.uint64(mask)) == np.uint64(value)]))
        got = st.masked_norm2(mask, value)
        assert abs(got - want) < 4 * TOL[dtype], (mask, value, got, want)
    with pytest.raises(_lib.EngineError):
        st.masked_norm2(1, 2)  # value outside the mask


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [14, 17, 21])
def test_expect_single_flip_vs_oracle(dtype, n):
    """tcb200_expect_single_flip (X_j / Y_j with Z dressing, register-pair kernel) through
    DeviceState.expectation_terms: vs the oracle's closed form and vs tcb200_expect_pauli"""
    rng = np.random.default_rng(n)
    psi = _rand_state(rng, n)
    st = DeviceState(n, dtype)
    st.load(psi)
    terms = []
    for j in range(n):  # every flip bit, both types, random Z dressing
        q = n - 1 - j
        others = [k for k in range(n) if k != q]
        terms.append(([q], [], sorted(rng.choice(others, size=int(rng.integers(0, 4)), replace=False).tolist())))
        terms.append(([], [q], sorted(rng.choice(others, size=int(rng.integers(0, 3)), replace=False).tolist())))
    fl, sg, ny = [], [], []
    for x, y, z in terms:
        f, s, k = orc.pauli_masks(n, x, y, z)
        fl.append(f)
        sg.append(s)
        ny.append(k)
    before = engine.STATS["expect_launches"]
    got = st.expectation_terms(fl, sg, ny)[0]
    launches = engine.STATS["expect_launches"] - before
    assert launches <= (2 * n + 11) // 12 + 3  # 12 strings per launch, not 8
    want = np.array([orc.pauli_expectation(psi, n, x, y, z) for x, y, z in terms])
    tol = 2e-6 if dtype == "complex64" else 1e-12
    assert np.max(np.abs(got - want)) < tol, np.max(np.abs(got - want))
    old = DeviceState.use_single_flip
    try:
        DeviceState.use_single_flip = False
        ref = st.expectation_terms(fl, sg, ny)[0]
    finally:
        DeviceState.use_single_flip = old
    assert np.max(np.abs(got - ref)) < tol


def test_expect_single_flip_batched_rows():
    n, B = 15, 3
    rng = np.random.default_rng(4)
    rows = [_rand_state(rng, n) for _ in range(B)]
    st = DeviceState(n, "complex64", batch=B)
    st.buf.copy_(torch.from_numpy(np.stack(rows)).to(st.device, dtype=torch.complex64))
    terms = [([n - 1 - j], [], [(n - j) % n] if (n - j) % n != n - 1 - j else []) for j in range(n)]
    fl, sg, ny = zip(*[orc.pauli_masks(n, x, y, z) for x, y, z in terms])
    got = st.expectation_terms(list(fl), list(sg), list(ny))
    for b in range(B):
        want = np.array([orc.pauli_expectation(rows[b], n, x, y, z) for x, y, z in terms])
        assert np.max(np.abs(got[b] - want)) < 2e-6
