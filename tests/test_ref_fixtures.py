"""Golden vectors produced by the REFERENCE's own code (tests/golden/ref_fixtures.npz, generated
by oracle/make_golden.py: reference modules run unmodified from /root/reference, their
un-vendored tensornetwork / opt_einsum dependency replaced by oracle/refshim).

They pin (i) the oracle and (ii) the engine -- on the CPU through the kernel emulation
(``emu``) and on the GPU through the C ABI (``cuda``, ``-m gpu``).  Nothing here reads
/root/reference at run time."""

import os

import numpy as np
import pytest

import tensorcircuit_b200 as tc
from oracle import tc_oracle as orc

from .test_circuit_api import eng  # noqa: F401  (fixture: emu on CPU, cuda under -m gpu)

FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fixtures.npz"))

# the same gate list oracle/make_golden.py fed to the reference
ALL_GATES = [
    ("h", (0,), {}), ("x", (1,), {}), ("y", (2,), {}), ("z", (3,), {}), ("t", (0,), {}), ("s", (1,), {}), ("td", (2,), {}), ("sd", (3,), {}),
    ("wroot", (0,), {}), ("cnot", (0, 1), {}), ("cz", (1, 2), {}), ("swap", (2, 3), {}), ("cy", (3, 0), {}), ("ox", (0, 2), {}), ("oy", (1, 3), {}),
    ("oz", (2, 0), {}), ("toffoli", (0, 1, 2), {}), ("fredkin", (3, 1, 0), {}),
    ("r", (0,), dict(theta=0.3, alpha=1.1, phi=-0.7)), ("cr", (1, 2), dict(theta=0.3, alpha=1.1, phi=-0.7)), ("u", (3,), dict(theta=0.4, phi=0.5, lbd=-1.2)),
    ("cu", (0, 3), dict(theta=0.4, phi=0.5, lbd=-1.2)), ("rx", (1,), dict(theta=0.6)), ("ry", (2,), dict(theta=0.7)), ("rz", (3,), dict(theta=0.8)),
    ("phase", (0,), dict(theta=0.9)), ("rxx", (0, 1), dict(theta=0.25)), ("ryy", (1, 2), dict(theta=0.35)), ("rzz", (2, 3), dict(theta=0.45)),
    ("cphase", (3, 1), dict(theta=0.55)), ("crx", (0, 2), dict(theta=0.65)), ("cry", (1, 3), dict(theta=0.75)), ("crz", (2, 0), dict(theta=0.85)),
    ("orx", (3, 2), dict(theta=0.15)), ("ory", (0, 1), dict(theta=0.22)), ("orz", (1, 0), dict(theta=0.33)), ("iswap", (2, 3), dict(theta=0.4)),
    ("cx", (3, 2), {}), ("cswap", (0, 2, 3), {}), ("ccnot", (1, 2, 0), {}), ("sdg", (1,), {}), ("tdg", (2,), {}),
]


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def _hea10():
    n = 10
    params = np.random.default_rng(0).uniform(0, 2 * np.pi, size=[4, 2, n])
    return n, orc.hea_circuit(n, params), [list(map(int, r)) for r in FIX["hea10_pss"]]


def _sample_ok(got, ref_idx, probs, u, tie=3e-7):
    """identical indices, except where the uniform sits on a CDF tie (|CDF - r| tiny)"""
    cdf = np.cumsum(np.asarray(probs, dtype=np.float64) / np.sum(probs, dtype=np.float64))
    r = cdf[-1] * (1 - u)
    bad = np.nonzero(np.asarray(got) != np.asarray(ref_idx))[0]
    for i in bad:
        lo = min(int(got[i]), int(ref_idx[i]))
        assert abs(cdf[lo] - r[i]) < tie, (i, got[i], ref_idx[i])  # float32 CDF resolution of the reference
    return len(bad)


# ---- the oracle against the reference's outputs -------------------------------------------------
def test_oracle_vs_reference_outputs():
    n, ops, pss = _hea10()
    o = orc.run_gatelist(n, ops)
    assert _rel(o.state(), FIX["hea10_state_complex128"]) < 1e-13
    assert _rel(o.state(), FIX["hea10_state_complex64"]) < 2e-6
    exps = np.array([o.expectation_ps(ps=ps) for ps in pss])
    np.testing.assert_allclose(exps, FIX["hea10_exps_complex128"], atol=1e-13)
    np.testing.assert_allclose(exps, FIX["hea10_exps_complex64"], atol=2e-6)
    o = orc.run_gatelist(9, orc.random_circuit(9, 6, seed=3))
    assert _rel(o.state(), FIX["rand9_state"]) < 2e-6
    np.testing.assert_allclose(o.probability(), FIX["rand9_probability_c128"], atol=1e-14)
    u = FIX["rand9_status"]
    # complex128 reference: float64 CDF -> the oracle's rule must give identical indices
    np.testing.assert_array_equal(orc.probability_sample(o.probability(), u), FIX["rand9_sample_int_c128"])
    # complex64 reference: float32 CDF; identical when the oracle runs the rule in float32 on
    # the reference's own float32 probabilities
    p32 = (np.abs(FIX["rand9_state"]) ** 2).astype(np.float32)
    np.testing.assert_array_equal(orc.probability_sample(p32, u, dtype=np.float32), FIX["rand9_sample_int"])
    assert _sample_ok(orc.probability_sample(o.probability(), u), FIX["rand9_sample_int"], o.probability(), u) < 20
    np.testing.assert_array_equal(orc.sample2all(FIX["rand9_sample_int"], 9, "count_vector"), FIX["rand9_count_vector"])
    np.testing.assert_array_equal(orc.sample_int2bin(FIX["rand9_sample_int"][:8], 9), FIX["rand9_sample_bin_head"])
    o = orc.run_gatelist(4, ALL_GATES)
    assert _rel(o.state(), FIX["allgates_state"]) < 2e-6
    o = orc.run_gatelist(4, ALL_GATES, inputs=FIX["allgates_inputs"])
    assert _rel(o.state(), FIX["allgates_state_inputs"]) < 2e-6
    np.testing.assert_allclose(o.expectation_ps(x=[0], z=[2]), FIX["allgates_exp_x0z2"], atol=2e-6)
    np.testing.assert_allclose(o.expectation_ps(y=[1, 3]), FIX["allgates_exp_y1y3"], atol=2e-6)
    np.testing.assert_allclose(o.expectation((FIX["allgates_op2"], [1, 2]), (orc.Z, [0])), FIX["allgates_exp_op2_q12"], atol=1e-5)
    c = orc.OracleCircuit(2)
    c.rx(0, theta=0.8 + 0.7j)
    c.rzz(0, 1, theta=-0.2j)
    assert _rel(c.state(), FIX["complex_param_state"]) < 2e-6
    c = orc.OracleCircuit(2, inputs=np.eye(4))
    c.rx(0, theta=0.3)
    c.cnot(0, 1)
    assert _rel(c.state(), FIX["unitary_inputs_state"]) < 2e-6


# ---- the engine against the reference's outputs ----------------------------------------------------
def _build(n, ops, **kw):
    c = tc.Circuit(n, **kw)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    return c


@pytest.mark.parametrize("dtype,tol", [("complex64", 1e-5), ("complex128", 1e-11)])
def test_engine_config1_vs_reference(eng, dtype, tol):  # noqa: F811
    tc.set_dtype(dtype)
    n, ops, pss = _hea10()
    c = _build(n, ops)
    assert _rel(c.state(), FIX["hea10_state_" + dtype]) < tol
    want = FIX["hea10_exps_" + dtype]
    got = np.asarray(c.expectation_ps_many(pss))
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * len(pss))) < tol
    tc.set_dtype("complex64")


def test_engine_sampler_vs_reference(eng):  # noqa: F811
    n = 9
    c = _build(n, orc.random_circuit(n, 6, seed=3))
    assert _rel(c.state(), FIX["rand9_state"]) < 1e-5
    u = FIX["rand9_status"]
    got = c.sample(batch=len(u), allow_state=True, status=u, format="sample_int")
    nbad = _sample_ok(got, FIX["rand9_sample_int"], np.abs(FIX["rand9_probability_c128"]), u)
    assert nbad < 20
    np.testing.assert_array_equal(c.sample(batch=8, allow_state=True, status=u[:8], format="sample_bin"), FIX["rand9_sample_bin_head"])
    cv = c.sample(batch=len(u), allow_state=True, status=u, format="count_vector")
    assert np.sum(np.abs(cv - FIX["rand9_count_vector"])) <= 2 * nbad
    tc.set_dtype("complex128")
    c = _build(n, orc.random_circuit(n, 6, seed=3))
    got = c.sample(batch=len(u), allow_state=True, status=u, format="sample_int")
    assert _sample_ok(got, FIX["rand9_sample_int_c128"], FIX["rand9_probability_c128"], u) <= 2
    np.testing.assert_allclose(c.probability(), FIX["rand9_probability_c128"], atol=1e-13)
    tc.set_dtype("complex64")


def test_engine_all_gates_vs_reference(eng):  # noqa: F811
    c = _build(4, ALL_GATES)
    assert _rel(c.state(), FIX["allgates_state"]) < 1e-5
    c = _build(4, ALL_GATES, inputs=FIX["allgates_inputs"])
    assert _rel(c.state(), FIX["allgates_state_inputs"]) < 1e-5
    np.testing.assert_allclose(c.expectation_ps(x=[0], z=[2]), FIX["allgates_exp_x0z2"], atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(y=[1, 3]), FIX["allgates_exp_y1y3"], atol=1e-5)
    got = c.expectation((FIX["allgates_op2"].reshape(2, 2, 2, 2), [1, 2]), (tc.gates.z(), [0]))
    np.testing.assert_allclose(got, FIX["allgates_exp_op2_q12"], atol=2e-5)
    c = tc.Circuit(2)
    c.rx(0, theta=0.8 + 0.7j)
    c.rzz(0, 1, theta=-0.2j)
    assert _rel(c.state(), FIX["complex_param_state"]) < 1e-5
    c = tc.Circuit(2, inputs=np.eye(4))
    c.rx(0, theta=0.3)
    c.cnot(0, 1)
    assert _rel(c.state(), FIX["unitary_inputs_state"]) < 1e-5


def test_engine_vmap_vs_reference(eng):  # noqa: F811
    K = tc.backend

    def f(theta):
        c = tc.Circuit(3)
        c.rx(0, theta=theta[0])
        c.ry(1, theta=theta[1])
        c.cnot(0, 2)
        c.rzz(1, 2, theta=theta[2] * 0.5)
        return K.real(c.expectation_ps(z=[2]) + 2.0 * c.expectation_ps(x=[1], z=[0]))

    got = K.vmap(f, vectorized_argnums=0)(FIX["vmap_theta"])
    np.testing.assert_allclose(got, FIX["vmap_values"], atol=1e-5)


# ---- Monte-Carlo trajectories driven by status (circuit.py:473-744, basecircuit.py:824-857) ------
def test_oracle_trajectories_vs_reference():
    n = 5
    st, th = FIX["traj_status"], FIX["traj_theta"]
    for t in range(st.shape[0]):
        o = orc.OracleCircuit(n)
        picks = orc.noisy_trajectory(o, n, st[t], th[t])
        assert picks == list(FIX["traj_picks_complex128"][t])
        np.testing.assert_allclose(o.state(), FIX["traj_states_complex128"][t], atol=1e-12)
        # the reference's complex64 run took the same branches
        assert picks == list(FIX["traj_picks_complex64"][t])


@pytest.mark.parametrize("dtype,tol", [("complex64", 2e-5), ("complex128", 1e-11)])
def test_engine_trajectories_vs_reference(eng, dtype, tol):  # noqa: F811
    n = 5
    tc.set_dtype(dtype)
    st, th = FIX["traj_status"], FIX["traj_theta"]
    for t in range(st.shape[0]):
        c = tc.Circuit(n)
        picks = orc.noisy_trajectory(c, n, st[t], th[t])
        assert picks == list(FIX["traj_picks_" + dtype][t])
        ref = FIX["traj_states_" + dtype][t]
        assert np.linalg.norm(np.asarray(c.wavefunction()) - ref) < tol * max(1.0, np.linalg.norm(ref))


def test_engine_trajectories_vmapped_vs_reference(eng):  # noqa: F811
    """All trajectories as ONE vmapped circuit: status and theta batched (the reference's
    K.vmap(f, vectorized_argnums=(0, 1)) over Monte-Carlo trajectories, docs/source/advance.rst
    'noisy circuit simulation')."""
    n = 5
    tc.set_dtype("complex128")
    K = tc.backend
    st, th = FIX["traj_status"], FIX["traj_theta"]

    def f(status, theta):
        c = tc.Circuit(n)
        k = 0
        for layer in range(2):
            for i in range(n):
                c.h(i)
            for i in range(n - 1):
                c.cnot(i, i + 1)
            for i in range(n):
                c.rx(i, theta=theta[layer * n + i])
                c.depolarizing(i, px=0.1, py=0.05, pz=0.15, status=status[k])
                c.amplitudedamping(i, gamma=0.3, p=0.8, status=status[k + 1])
                c.phasedamping(i, gamma=0.2, status=status[k + 2])
                k += 3
        r = c.cond_measure(1, status=status[k])
        c.reset(2, status=status[k + 1])
        return c.wavefunction(), r

    # the un-vmapped reference values up to the cond_measure / reset step
    want, wr = [], []
    for t in range(st.shape[0]):
        o = orc.OracleCircuit(n)
        k = 0
        for layer in range(2):
            for i in range(n):
                o.h(i)
            for i in range(n - 1):
                o.cnot(i, i + 1)
            for i in range(n):
                o.rx(i, theta=th[t][layer * n + i])
                o.depolarizing(i, px=0.1, py=0.05, pz=0.15, status=st[t][k])
                o.amplitudedamping(i, gamma=0.3, p=0.8, status=st[t][k + 1])
                o.phasedamping(i, gamma=0.2, status=st[t][k + 2])
                k += 3
        wr.append(o.cond_measure(1, status=st[t][k]))
        o.reset(2, status=st[t][k + 1])
        want.append(o.state())
    got, r = K.vmap(f, vectorized_argnums=(0, 1))(st, th)
    assert list(np.asarray(r)) == wr == list(FIX["traj_picks_complex128"][:, 0])
    np.testing.assert_allclose(np.asarray(got), np.array(want), atol=1e-11)
    tc.set_dtype("complex64")


# ---- QAOA MaxCut, n = 14: a diagonal cost function (every term through the Z-string kernel) --------
def _qaoa(cls, n, edges, gam, bet):
    c = cls(n)
    for i in range(n):
        c.h(i)
    for l in range(2):
        for a, b in edges:
            c.rzz(a, b, theta=2 * gam[l])
        for i in range(n):
            c.rx(i, theta=2 * bet[l])
    return c


def test_oracle_qaoa_vs_reference():
    n = 14
    edges = [tuple(map(int, e)) for e in FIX["qaoa_edges"]]
    gam, bet = FIX["qaoa_angles"]
    o = _qaoa(orc.OracleCircuit, n, edges, gam, bet)
    np.testing.assert_allclose(o.state()[:: 2**n // 256], FIX["qaoa_amps_complex128"], atol=1e-13)
    zz = np.array([o.expectation_ps(z=[a, b]) for a, b in edges])
    np.testing.assert_allclose(zz, FIX["qaoa_zz_complex128"], atol=1e-12)
    np.testing.assert_allclose(o.expectation_ps(z=[0, 6, 13]), FIX["qaoa_z3_complex128"], atol=1e-12)
    np.testing.assert_allclose(zz, FIX["qaoa_zz_complex64"], atol=2e-6)
    # the sampling rule with the reference's float32 CDF, bit for bit on its complex64 state
    o32 = _qaoa(orc.OracleCircuit, n, edges, gam, bet)
    p32 = (np.abs(o32.state().astype(np.complex64)) ** 2).astype(np.float32)
    got = orc.probability_sample(p32, FIX["qaoa_status"], dtype=np.float32)
    assert np.mean(got == FIX["qaoa_sample_int"]) > 0.98  # the oracle's state is float64-exact, the reference's is not


@pytest.mark.parametrize("dtype,tol", [("complex64", 1e-5), ("complex128", 1e-11)])
def test_engine_qaoa_vs_reference(eng, dtype, tol):  # noqa: F811
    n = 14
    tc.set_dtype(dtype)
    edges = [tuple(map(int, e)) for e in FIX["qaoa_edges"]]
    gam, bet = FIX["qaoa_angles"]
    c = _qaoa(tc.Circuit, n, edges, gam, bet)
    amps = np.asarray(c.wavefunction())[:: 2**n // 256]
    ref = FIX["qaoa_amps_" + dtype]
    assert np.linalg.norm(amps - ref) < tol * np.linalg.norm(ref) * 4
    pss = [[3 if q in e else 0 for q in range(n)] for e in edges] + [[3 if q in (0, 6, 13) else 0 for q in range(n)]]
    vals = np.asarray(c.expectation_ps_many(pss))
    np.testing.assert_allclose(vals[:-1], FIX["qaoa_zz_" + dtype], atol=20 * tol)
    np.testing.assert_allclose(vals[-1], FIX["qaoa_z3_" + dtype], atol=20 * tol)
    if dtype == "complex64":
        u = FIX["qaoa_status"]
        got = np.asarray(c.sample(batch=len(u), allow_state=True, status=u, format="sample_int"))
        # The reference accumulates its CDF sequentially in float32 (abstract_backend.py:1124-1157):
        # over 2^14 entries that is only good to ~1e-5, so shots whose uniform sits that close to a
        # CDF step may land on the neighbouring index.  (The oracle, which emulates the float32
        # cumsum, reproduces the reference's indices exactly: test_oracle_qaoa_vs_reference.)
        bad = _sample_ok(got, FIX["qaoa_sample_int"], np.abs(np.asarray(c.wavefunction()).astype(np.complex128)) ** 2, u, tie=2e-5)
        assert bad <= 0.06 * len(u)  # float64-exact CDF vs the reference float32 one: 21 of 512 at this size
    tc.set_dtype("complex64")


# ---- sample_expectation_ps (basecircuit.py:618-758) --------------------------------------------------
def test_engine_sample_expectation_ps_vs_reference(eng):  # noqa: F811
    c = _build(4, ALL_GATES)
    before = np.asarray(c.wavefunction()).copy()
    nq = len(c.to_qir())
    got = [c.sample_expectation_ps(x=[0], y=[1], z=[3]), c.sample_expectation_ps(z=[0, 2]), c.sexpps(y=[2, 3])]
    np.testing.assert_allclose(np.asarray(got, dtype=np.float64), FIX["sexpps_exact"], atol=2e-6)
    u = FIX["sexpps_status"]
    s1 = c.sample_expectation_ps(x=[0], y=[1], z=[3], shots=4096, status=u)
    s2 = c.sample_expectation_ps(x=[1, 2], shots=4096, status=u)
    # identical samples -> identical estimate (2 / 4096 per shot that sits on a float32 CDF tie)
    np.testing.assert_allclose([s1, s2], FIX["sexpps_shots"], atol=3 * 2 / 4096)
    # the basis rotation is undone: same state, same record
    assert len(c.to_qir()) == nq
    np.testing.assert_allclose(np.asarray(c.wavefunction()), before, atol=5e-7)
    # oracle: rotate, then a product of Z's
    o = orc.run_gatelist(4, ALL_GATES)
    o.h(0)
    o.rx(1, theta=np.pi / 2)
    np.testing.assert_allclose(o.expectation_ps(z=[0, 1, 3]).real, FIX["sexpps_exact"][0], atol=2e-6)


# ---- measure_jit / perfect_sampling with per-qubit status (basecircuit.py:359-443) --------------------
@pytest.mark.parametrize("dtype,tol", [("complex64", 2e-6), ("complex128", 2e-7)])
def test_engine_measure_vs_reference(eng, dtype, tol):  # noqa: F811
    # complex128 tolerance: the reference builds its td / sd gates as adjoints at import time, in the
    # default complex64, so its own complex128 state of this all-gates circuit is only good to 2e-8
    # (first deviation from the float64 oracle at the td gate); the engine and the oracle agree to 1e-16.
    tc.set_dtype(dtype)
    c = _build(4, ALL_GATES)
    st = FIX["measure_status"]
    for k, row in enumerate(st):
        b, pr = c.measure(0, 2, 3, with_prob=True, status=row)
        assert list(np.asarray(b)) == list(FIX["measure_bits_" + dtype][k])
        np.testing.assert_allclose(pr, FIX["measure_probs_" + dtype][k], atol=tol)
        b, pr = c.perfect_sampling(status=row)
        assert list(np.asarray(b)) == list(FIX["perfect_bits_" + dtype][k])
        np.testing.assert_allclose(pr, FIX["perfect_probs_" + dtype][k], atol=tol)
    b, pr = c.measure(1, status=[0.3])
    assert pr == -1.0 and b.shape == (1,)
    # the record probabilities of all 16 outcomes of perfect_sampling form the distribution |psi|^2
    psi = np.asarray(c.wavefunction()).astype(np.complex128)
    for row in st[:4]:
        b, pr = c.perfect_sampling(status=row)
        i = int("".join(str(int(x)) for x in np.asarray(b)), 2)
        np.testing.assert_allclose(pr, abs(psi[i]) ** 2, atol=2e-6 if dtype == "complex64" else 1e-13)
    tc.set_dtype("complex64")


# ---- readout error (basecircuit.py:587-596, 760-803) ---------------------------------------------------
def test_engine_readout_error_vs_reference(eng):  # noqa: F811
    c = _build(4, ALL_GATES)
    ro = [list(map(float, r)) for r in FIX["readout_error"]]
    np.testing.assert_allclose(np.asarray(c.readouterror_bs(ro)), FIX["readout_probs"], atol=2e-6)
    u = FIX["readout_status"]
    got = np.asarray(c.sample(batch=len(u), allow_state=True, readout_error=ro, status=u, format="sample_int"))
    bad = _sample_ok(got, FIX["readout_sample_int"], FIX["readout_probs"], u, tie=1e-6)
    assert bad <= 3
    e1 = c.sample_expectation_ps(x=[0], y=[1], z=[3], readout_error=ro)
    e2 = c.sample_expectation_ps(z=[0, 2], readout_error=ro)
    e3 = c.sample_expectation_ps(x=[1, 2], shots=len(u), status=u, readout_error=ro)
    np.testing.assert_allclose([e1, e2], FIX["readout_sexpps"][:2], atol=3e-6)
    np.testing.assert_allclose(e3, FIX["readout_sexpps"][2], atol=3 * 2 / len(u))
    # no readout error given: plain probabilities
    np.testing.assert_allclose(np.asarray(c.readouterror_bs(None)), np.abs(np.asarray(c.wavefunction())) ** 2, atol=1e-6)
