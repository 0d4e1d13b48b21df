"""Parity of the ``tc.Circuit`` surface with the reference.

Test bodies follow the reference's own tests (cited per test, paths under /root/reference/)
and compare against the oracle.  Each body runs twice:
  * ``emu``  -- CPU: host logic + the kernel bodies through tests/emu (``-m "not gpu"``),
  * ``cuda`` -- the real sm_100a kernels through the C ABI (``-m gpu``)."""

import numpy as np
import pytest

import tensorcircuit_b200 as tc
from oracle import tc_oracle as orc
from oracle.tc_oracle import OracleCircuit


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def eng(request, monkeypatch):
    if request.param == "emu":
        from .fake_state import FakeCOO, FakeState

        monkeypatch.setattr(tc.engine, "DeviceState", FakeState)
        monkeypatch.setattr(tc.engine, "DeviceCOO", FakeCOO)
    tc.set_dtype("complex64")
    yield request.param
    tc.set_dtype("complex64")


@pytest.fixture
def highp(eng):
    tc.set_dtype("complex128")
    yield
    tc.set_dtype("complex64")


def A(x):
    return np.asarray(x)


# ---- tests/test_circuit.py ---------------------------------------------------------------------
def test_wavefunction(eng):
    # tests/test_circuit.py:22-43
    qc = tc.Circuit(2)
    qc.unitary(0, 1, unitary=tc.gates.Gate(np.arange(16).reshape(2, 2, 2, 2).astype(np.complex64)))
    assert np.real(qc.wavefunction()[2]) == 8
    qc = tc.Circuit(2)
    qc.unitary(1, 0, unitary=tc.gates.Gate(np.arange(16).reshape(2, 2, 2, 2).astype(np.complex64)))
    assert np.real(qc.wavefunction()[2]) == 4
    qc = tc.Circuit(2)
    qc.unitary(0, unitary=tc.gates.Gate(np.arange(4).reshape(2, 2).astype(np.complex64)))
    assert np.real(qc.wavefunction()[2]) == 2


def test_basics(eng):
    # tests/test_circuit.py:47-52
    c = tc.Circuit(2)
    c.x(0)
    np.testing.assert_allclose(c.amplitude("10"), 1.0)
    c.CNOT(0, 1)
    np.testing.assert_allclose(c.amplitude("11"), 1.0)


def test_gates_in_circuit(eng):
    # tests/test_circuit.py:64-68
    c = tc.Circuit(2, inputs=np.eye(2**2))
    c.iswap(0, 1)
    ans = A(tc.gates.iswap_gate().tensor).reshape([4, 4])
    np.testing.assert_allclose(A(c.state()).reshape([4, 4]), ans, atol=1e-5)


def test_control_vgate(eng):
    # tests/test_circuit.py:71-77
    c = tc.Circuit(2)
    c.x(1)
    c.crx(1, 0, theta=0.3)
    np.testing.assert_allclose(c.expectation([tc.gates._z_matrix, 0]), 0.95533645, atol=1e-5)


def test_adjoint_gate_circuit(eng):
    # tests/test_circuit.py:80-84
    c = tc.Circuit(1)
    c.X(0)
    c.SD(0)
    np.testing.assert_allclose(c.state(), np.array([0.0, -1.0j]))


def test_expectation(eng):
    # tests/test_circuit.py:240-245
    c = tc.Circuit(2)
    c.H(0)
    np.testing.assert_allclose(c.expectation((tc.gates.z(), [0])), 0, atol=1e-7)


def test_exp1(eng):
    # tests/test_circuit.py:248-272 : exp and exp1 agree for an involutory generator
    zz = np.kron(tc.gates._z_matrix, tc.gates._z_matrix)
    c = tc.Circuit(2)
    c.H(0)
    c.H(1)
    c.exp1(0, 1, unitary=zz, theta=0.35)
    c2 = tc.Circuit(2)
    c2.H(0)
    c2.H(1)
    c2.exp(0, 1, unitary=zz, theta=0.35)
    np.testing.assert_allclose(c.state(), c2.state(), atol=1e-6)


def test_complex128(highp):
    # tests/test_circuit.py:275-280
    c = tc.Circuit(2)
    c.H(1)
    c.rx(0, theta=1.0j)
    c.wavefunction()
    assert A(c.wavefunction()).dtype == np.complex128
    np.testing.assert_allclose(c.expectation((tc.gates.z(), [1])), 0, atol=1e-9)


def test_single_qubit(eng):
    # tests/test_circuit.py:319-323
    c = tc.Circuit(1)
    c.H(0)
    np.testing.assert_allclose(c.state(), np.array([1, 1]) / np.sqrt(2), atol=1e-4)


def test_complex_parameter(eng):
    # tests/test_circuit.py:343-352 : complex gate parameters (non-unitary matrices)
    c = tc.Circuit(2)
    c.rx(0, theta=0.8 + 0.7j)
    c.rzz(0, 1, theta=-0.2j)
    o = OracleCircuit(2)
    o.rx(0, theta=0.8 + 0.7j)
    o.rzz(0, 1, theta=-0.2j)
    np.testing.assert_allclose(c.state(), o.state(), atol=1e-5)


def test_unitary(eng):
    # tests/test_circuit.py:404-412
    c = tc.Circuit(2, inputs=np.eye(4))
    c.X(0)
    c.Y(1)
    answer = np.kron(A(tc.gates.x().tensor), A(tc.gates.y().tensor))
    np.testing.assert_allclose(A(c.wavefunction()).reshape([4, 4]), answer, atol=1e-4)


def test_expectation_ps(eng):
    # tests/test_circuit.py:416-429
    c = tc.Circuit(2)
    c.X(0)
    np.testing.assert_allclose(c.expectation_ps(z=[0, 1]), -1, atol=1e-5)
    c = tc.Circuit(2)
    c.H(0)
    np.testing.assert_allclose(c.expectation_ps(z=[1], x=[0]), 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(ps=[1, 3]), 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(z=[1, 2], ps=[1, 3]), 1, atol=1e-5)


def test_probability(eng):
    # tests/test_circuit.py:432-443
    c = tc.Circuit(2)
    c.h(0)
    c.h(1)
    np.testing.assert_allclose(c.probability(), np.ones(4) / 4, atol=1e-5)


def test_mixed_measurement_circuit(eng):
    # tests/test_circuit.py:500-554 : <X_i>
    n = 4
    c = tc.Circuit(n)
    for i in range(n):
        c.H(i)
    for j in range(2):
        for i in range(n):
            c.cnot(i, (i + 1) % n)
        for i in range(n):
            c.rz(i, theta=1.0)
    v = [np.real(c.expectation_ps(x=[i])) for i in range(n)]
    np.testing.assert_allclose(v, [0.157729, 0.157729, 0.157728, 0.085221], atol=1e-5)
    # same through the structure-tensor measurement (onehot weights), as the reference test does
    for i in range(n):
        s = np.zeros((n, 4))
        s[:, 0] = 1
        s[i] = [0, 1, 0, 0]
        r = tc.templates.measurements.parameterized_measurements(c, s, onehot=False)
        np.testing.assert_allclose(r, v[i], atol=1e-5)

    # the reference's own body: vvag over the structures, value and gradient pinned (:536-554)
    def f(params, structures):
        c = tc.Circuit(n)
        for i in range(n):
            c.H(i)
        for j in range(2):
            for i in range(n):
                c.cnot(i, (i + 1) % n)
            for i in range(n):
                c.rz(i, theta=params[j, i])
        obs = []
        for i in range(n):
            obs.append([tc.gates.Gate(sum([structures[i, k] * g.tensor for k, g in enumerate(tc.gates.pauli_gates)])), (i,)])
        return tc.backend.real(c.expectation(*obs, reuse=False))

    structures = np.eye(4)[np.eye(n, dtype=int)]  # onehot(eye(n), 4): [n, n, 4]
    v, g = tc.backend.vvag(f, vectorized_argnums=1, argnums=0)(np.ones([2, n]), structures)
    np.testing.assert_allclose(v, [0.157729, 0.157729, 0.157728, 0.085221], atol=1e-5)
    np.testing.assert_allclose(g[0], [-0.378372, -0.624019, -0.491295, -0.378372], atol=1e-5)


def test_circuit_replace_inputs(eng):
    # tests/test_circuit.py:571-580
    n = 3
    c = tc.Circuit(n, inputs=np.zeros([2**n]))
    for i in range(n):
        c.H(i)
    evenstate = np.ones([2**n])
    evenstate /= np.linalg.norm(evenstate)
    c.replace_inputs(evenstate)
    for i in range(n):
        np.testing.assert_allclose(c.expectation_ps(z=[i]), 1.0, atol=1e-5)


def test_toqir(eng):
    # tests/test_circuit.py:707-727
    c = tc.Circuit(3)
    c.H(0)
    c.rx(1, theta=tc.array_to_tensor(0.7))
    c.exp1(0, 1, unitary=tc.gates._zz_matrix, theta=tc.array_to_tensor(-0.2))
    z1 = c.expectation((tc.gates.z(), [1]))
    qirs = c.to_qir()
    c = tc.Circuit.from_qir(qirs, circuit_params={"nqubits": 3})
    z2 = c.expectation((tc.gates.z(), [1]))
    np.testing.assert_allclose(z1, z2, atol=1e-5)
    c.append_from_qir(qirs)
    z3 = c.expectation((tc.gates.z(), [1]))
    np.testing.assert_allclose(z3, 0.202728, atol=1e-5)
    assert qirs[1]["name"] == "rx" and qirs[1]["index"] == (1,)


def test_circuit_append(eng):
    # tests/test_circuit.py:808-818
    c = tc.Circuit(2)
    c1 = tc.Circuit(1)
    c1.x(0)
    c.append(c1, [1])
    np.testing.assert_allclose(c.state(), np.array([0, 1, 0, 0]), atol=1e-5)


def test_expectation_y_bug(eng):
    # tests/test_circuit.py:1278-1281
    c = tc.Circuit(1, inputs=1 / np.sqrt(2) * np.array([-1, 1.0j]))
    np.testing.assert_allclose(c.expectation_ps(y=[0]), -1, atol=1e-5)


def test_circuit_inverse(eng):
    # tests/test_circuit.py:1314-1343
    rng = np.random.default_rng(0)
    inputs = rng.uniform(size=[8])
    inputs /= np.linalg.norm(inputs)
    c = tc.Circuit(3, inputs=inputs)
    c.H(1)
    c.rx(0, theta=0.5)
    c.cnot(1, 2)
    c.rzz(0, 2, theta=-0.8)
    c.append(c.inverse())
    np.testing.assert_allclose(c.state(), inputs, atol=1e-5)
    c = tc.Circuit(3, inputs=inputs)
    c.iswap(0, 1)
    c.iswap(1, 0, theta=0.6)
    c.rxx(1, 2, theta=-0.2)
    c.cu(0, 1, lbd=2.0, theta=-0.7)
    c.r(2, alpha=0.3)
    c.sd(2)
    c.cx(1, 2)
    c.unitary(0, unitary=tc.gates._x_matrix)
    c.append(c.inverse())
    np.testing.assert_allclose(c.state(), inputs, atol=1e-5)


def test_minus_index(eng):
    # tests/test_circuit.py:1373-1380
    c = tc.Circuit(3)
    c.H(-2)
    c.H(0)
    np.testing.assert_allclose(np.real(c.expectation_ps(x=[0])), 1, atol=1e-5)
    np.testing.assert_allclose(np.real(c.expectation_ps(x=[1])), 1, atol=1e-5)
    np.testing.assert_allclose(np.real(c.expectation_ps(x=[-1])), 0, atol=1e-5)
    np.testing.assert_allclose(np.real(c.expectation_ps(z=[-2])), 0, atol=1e-5)


def test_errors(eng):
    c = tc.Circuit(3)
    with pytest.raises(ValueError, match="Cannot measure two operators in one index"):  # basecircuit.py:306-307
        c.expectation_ps(x=[0], z=[0])
    with pytest.raises(AssertionError):  # basecircuit.py:143
        c.cnot(1, 1)
    with pytest.raises(ValueError, match="Illegal index specification"):  # abstractcircuit.py:165
        c.rx("a", theta=0.1)
    with pytest.raises(AssertionError):  # circuit.py:92
        tc.Circuit(3, inputs=np.ones(4))
    with pytest.raises(ValueError, match="Unsupported data type"):  # cons.py:153
        tc.set_dtype("complex32")
    with pytest.raises(ValueError, match="unsupported format"):  # quantum.py:2366-2368
        c.sample(batch=2, allow_state=True, format="nonsense", status=[0.1, 0.2])


def test_list_index_broadcast(eng):
    # abstractcircuit.py:149-165 : c.rx([..], theta=[..])
    c = tc.Circuit(3)
    c.rx([0, 1, 2], theta=[0.1, 0.2, 0.3])
    c.rzz(range(2), range(1, 3), theta=0.4)
    o = OracleCircuit(3)
    o.rx([0, 1, 2], theta=[0.1, 0.2, 0.3])
    o.rzz(range(2), range(1, 3), theta=0.4)
    assert len(c.to_qir()) == 5
    np.testing.assert_allclose(c.state(), o.state(), atol=1e-6)


# ---- tests/test_gates.py -----------------------------------------------------------------------
def test_gate_matrices_match_oracle():
    for name in tc.Circuit.sgates:
        np.testing.assert_allclose(tc.gates.matrix_for_gate(getattr(tc.gates, name)(), tol=0), orc.gate_matrix(name), atol=1e-12, err_msg=name)
    p3 = dict(theta=0.3, alpha=1.1, phi=-0.7)
    cases = {
        "r": p3, "cr": p3, "u": dict(theta=0.3, phi=0.4, lbd=-1.2), "cu": dict(theta=0.3, phi=0.4, lbd=-1.2),
        "rx": dict(theta=0.3), "ry": dict(theta=0.3), "rz": dict(theta=0.3), "phase": dict(theta=0.3),
        "rxx": dict(theta=0.3), "ryy": dict(theta=0.3), "rzz": dict(theta=0.3), "cphase": dict(theta=0.3),
        "crx": dict(theta=0.3), "cry": dict(theta=0.3), "crz": dict(theta=0.3), "orx": dict(theta=0.3),
        "ory": dict(theta=0.3), "orz": dict(theta=0.3), "iswap": dict(theta=0.3),
        "exp": dict(unitary=np.kron(orc.X, orc.Y), theta=0.3), "exp1": dict(unitary=np.kron(orc.X, orc.Y), theta=0.3),
    }
    for name, p in cases.items():
        np.testing.assert_allclose(tc.gates.matrix_for_gate(getattr(tc.gates, name)(**p), tol=0), orc.gate_matrix(name, **p), atol=1e-12, err_msg=name)
    assert tc.Circuit.sgates == tc.circuit.sgates  # tests/test_gates.py:104-105


def test_gate_factories():
    # tests/test_gates.py:13-16
    np.testing.assert_almost_equal(tc.gates.r_gate(1, 2, 3).tensor, tc.gates.rgate_theoretical(1, 2, 3).tensor)
    # tests/test_gates.py:44-57 (ided)
    g = tc.gates.rx.ided()
    np.testing.assert_allclose(tc.backend.reshapem(g(theta=0.3).tensor), np.kron(np.eye(2), tc.gates.rx(theta=0.3).tensor), atol=1e-5)
    g1 = tc.gates.rx.ided(before=False)
    np.testing.assert_allclose(tc.backend.reshapem(g1(theta=0.3).tensor), np.kron(tc.gates.rx(theta=0.3).tensor, np.eye(2)), atol=1e-5)
    # tests/test_gates.py:100-105
    ans = np.array([[1.0, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1.0]])
    np.testing.assert_allclose(tc.gates.iswap_gate().tensor, ans.reshape([2, 2, 2, 2]), atol=1e-5)
    np.testing.assert_allclose(tc.gates.iswap_gate(theta=0).tensor, np.eye(4).reshape([2, 2, 2, 2]), atol=1e-5)
    # tests/test_gates.py:108-116
    ccx = tc.gates.x.controlled().controlled()
    assert ccx.n == "ccx" and ccx.ctrl == [1, 1]
    np.testing.assert_allclose(ccx().tensor, tc.backend.reshape2(tc.gates._toffoli_matrix))
    # tests/test_gates.py:137-141
    np.testing.assert_allclose(tc.gates.sd().tensor, tc.backend.adjoint(tc.gates._s_matrix))
    assert tc.gates.td.n == "td"
    # multicontrol (gates.py:868-942): dense form
    m = tc.gates.multicontrol_gate(tc.gates._zz_matrix, [1, 0, 1]).matrix()
    np.testing.assert_allclose(m, orc.multicontrol_matrix(np.kron(orc.Z, orc.Z), [1, 0, 1]))


def test_phase_cu_fsim_exp_any(eng):
    # tests/test_gates.py:18-22
    c = tc.Circuit(1)
    c.h(0)
    c.phase(0, theta=np.pi / 2)
    np.testing.assert_allclose(c.state()[1], 0.7071j, atol=1e-4)
    # tests/test_gates.py:25-31
    c = tc.Circuit(2)
    c.cu(0, 1, theta=np.pi / 2, phi=-np.pi / 4, lbd=np.pi / 4)
    m = c.matrix()
    np.testing.assert_allclose(m[2:, 2:], tc.gates._wroot_matrix, atol=1e-5)
    np.testing.assert_allclose(m[:2, :2], np.eye(2), atol=1e-5)
    # tests/test_gates.py:60-77
    c = tc.Circuit(2)
    c.iswap(0, 1, theta=-0.2)
    c.cphase(0, 1, theta=-0.3)
    ans = np.array([[1.0, 0, 0, 0], [0, 0.95105654, -0.309017j, 0], [0, -0.309017j, 0.95105654, 0], [0, 0, 0, 0.9553365 - 0.29552022j]])
    np.testing.assert_allclose(c.matrix(), ans, atol=1e-5)
    # tests/test_gates.py:80-91
    c = tc.Circuit(2)
    c.exp(0, 1, unitary=np.diag([1.0, -1, -1, 1]), theta=np.pi / 2)
    np.testing.assert_allclose(c.wavefunction()[0], -1j, atol=1e-6)
    # tests/test_gates.py:94-97
    c = tc.Circuit(2)
    c.any(0, unitary=np.eye(2))
    np.testing.assert_allclose(c.expectation((tc.gates.z(), [0])), 1.0)


def test_controlled(eng):
    # tests/test_gates.py:116-134
    ocx = tc.gates.x.controlled().ocontrolled()
    c = tc.Circuit(3)
    c.x(0)
    c.any(1, 0, 2, unitary=ocx())
    np.testing.assert_allclose(c.expectation([tc.gates.z(), [2]]), -1, atol=1e-5)
    crxgate = tc.gates.rx.controlled()
    c = tc.Circuit(2)
    c.x(0)
    tc.Circuit.crx_my = tc.Circuit.apply_general_variable_gate_delayed(crxgate)
    c.crx_my(0, 1, theta=0.3)
    np.testing.assert_allclose(c.expectation([tc.gates.z(), [1]]), 0.95533645, atol=1e-5)
    assert c.to_qir()[1]["name"] == "crx"


def test_rxx_gate(eng):
    # tests/test_gates.py:144-155
    c1 = tc.Circuit(3)
    c1.rxx(0, 1, theta=1.0)
    c1.ryy(0, 2, theta=0.5)
    c1.rzz(0, 1, theta=-0.5)
    c2 = tc.Circuit(3)
    c2.exp1(0, 1, theta=1.0 / 2, unitary=tc.gates._xx_matrix)
    c2.exp1(0, 2, theta=0.5 / 2, unitary=tc.gates._yy_matrix)
    c2.exp1(0, 1, theta=-0.5 / 2, unitary=tc.gates._zz_matrix)
    np.testing.assert_allclose(c1.state(), c2.state(), atol=1e-5)


# ---- templates ---------------------------------------------------------------------------------
def test_templates(eng):
    # tests/test_templates.py:17-26
    c = tc.Circuit(2)
    c.H(0)
    c.H(1)
    np.testing.assert_allclose(tc.templates.measurements.any_measurements(c, np.array([1, 1]), onehot=True), 1.0, atol=1e-5)
    np.testing.assert_allclose(tc.templates.measurements.any_measurements(c, np.array([3, 0]), onehot=True), 0.0, atol=1e-5)
    # tests/test_templates.py:29-39
    c = tc.Circuit(3)
    c.X(0)
    c.cnot(0, 1)
    c.H(-1)
    r = tc.templates.measurements.parameterized_local_measurements(c, structures=np.array([3, 3, 1]), onehot=True)
    np.testing.assert_allclose(r, np.array([-1, -1, 1]), atol=1e-5)
    # tests/test_templates.py:111-119
    f = tc.templates.blocks.state_centric(tc.templates.blocks.Bell_pair_block)
    s = f(np.array([1.0, 0, 0, 0]))
    np.testing.assert_allclose(s, np.array([0.0, 0.70710677, -0.70710677, 0]), atol=1e-5)


# ---- sampling ----------------------------------------------------------------------------------
def test_sample_formats(eng):
    # tests/test_circuit.py:1254-1275 / tests/test_quantum.py:467-489 (shapes), basecircuit.py:587-616
    n = 4
    c = tc.Circuit(n)
    for i in range(n):
        c.H(i)
    u = np.random.default_rng(0).random(9)
    r = c.sample(batch=9, allow_state=True, format="sample_bin", status=u)
    assert r.shape == (9, n)
    r2 = c.sample(batch=9, allow_state=True, format="sample_int", status=u)
    assert r2.shape == (9,)
    np.testing.assert_array_equal(tc.quantum.sample_bin2int(r, n), r2)
    cv = c.sample(batch=9, allow_state=True, format="count_vector", status=u)
    assert cv.shape == (2**n,) and cv.sum() == 9
    ct = c.sample(batch=9, allow_state=True, format="count_tuple", status=u)
    assert ct[1].sum() == 9
    assert sum(c.sample(batch=9, allow_state=True, format="count_dict_bin", status=u).values()) == 9
    assert sum(c.sample(batch=9, allow_state=True, format="count_dict_int", status=u).values()) == 9
    lst = c.sample(batch=3, allow_state=True, status=u[:3])
    assert len(lst) == 3 and lst[0][0].shape == (n,)
    np.testing.assert_allclose(lst[0][1], 1 / 16, atol=1e-6)
    one = c.sample(allow_state=True, status=u[:1])
    assert one[0].shape == (n,)
    # reference rule: same indices as the oracle for the same uniforms
    np.testing.assert_array_equal(r2, orc.probability_sample(np.ones(2**n), u))
    # no status: backend RNG
    tc.backend.set_random_state(42)
    assert c.sample(batch=5, allow_state=True, format="sample_int").shape == (5,)


def test_sample_expectation_consistency(eng):
    # tests/test_circuit.py:1383-1405
    c = tc.Circuit(3)
    c.H(0)
    c.cnot(0, 1)
    c.rx(2, theta=0.4)
    s = c.sample(batch=4096, allow_state=True, format="sample_bin", status=np.random.default_rng(3).random(4096))
    est = tc.quantum.correlation_from_samples([0, 1], s, 3)
    np.testing.assert_allclose(est, np.real(c.expectation_ps(z=[0, 1])), atol=5e-2)


def test_quantum_helpers():
    # tests/test_quantum.py:299-312, 456-464
    np.testing.assert_allclose(tc.quantum.spin_by_basis(2, 1), np.array([1, -1, 1, -1]))
    state = np.array([0.6, 0.4, 0, 0])
    np.testing.assert_allclose(tc.quantum.correlation_from_counts([0, 1], state), 0.2, atol=1e-6)
    np.testing.assert_allclose(tc.quantum.correlation_from_counts([1], state), 0.2, atol=1e-6)
    np.testing.assert_allclose(tc.quantum.correlation_from_samples([0, 1], np.array([0, 0, 3, 3, 3]), n=2), 1, atol=1e-5)
    x, y = tc.quantum.count_d2s(np.array([0.1, 0, -0.3, 0]))
    np.testing.assert_allclose(x, np.array([0, 2]))
    np.testing.assert_allclose(y, np.array([0.1, -0.3]))
    np.testing.assert_allclose(tc.quantum.count_s2d((x, y), 2), np.array([0.1, 0, -0.3, 0]))
    assert tc.quantum.ps2xyz([1, 2, 2, 0]) == {"x": [0], "y": [1, 2], "z": []}
    assert tc.quantum.xyz2ps({"x": [1], "z": [3]}, n=4) == [0, 1, 0, 3]
    # tests/test_backends.py:278-280
    np.testing.assert_allclose(tc.backend.searchsorted([-1, 3.3, 9.1, 10.0], np.array([0.0, 4.1, 12.0], dtype=np.float32)), [1, 2, 4])


# ---- configs of SURVEY 8(d) at oracle-checkable sizes -----------------------------------------
def _run_gatelist(n, ops, **kw):
    c = tc.Circuit(n, **kw)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    return c


@pytest.mark.parametrize("dtype,tol", [("complex64", 1e-5), ("complex128", 1e-11)])
def test_config1_hea10(eng, dtype, tol):
    """Config 1: 10-qubit HEA (rx/rzz/cnot, depth 4): wavefunction + expectation_ps."""
    tc.set_dtype(dtype)
    n = 10
    params = np.random.default_rng(0).uniform(0, 2 * np.pi, size=[4, 2, n])
    ops = orc.hea_circuit(n, params)
    assert len(ops) == 112
    c = _run_gatelist(n, ops)
    o = orc.run_gatelist(n, ops)
    psi = A(c.state())
    assert np.linalg.norm(psi - o.state()) / np.linalg.norm(o.state()) < tol
    terms = orc.tfim_terms(n)
    pss = [ps for _, ps in terms] + [list(r) for r in np.random.default_rng(0).integers(0, 4, size=[8, n])]
    want = np.array([o.expectation_ps(ps=ps) for ps in pss])
    got = A(c.expectation_ps_many(pss))
    scale = np.maximum(np.abs(want), 1e-3 * len(pss))
    assert np.max(np.abs(got - want) / scale) < tol
    one = c.expectation_ps(ps=pss[-1])
    assert abs(one - want[-1]) / scale[-1] < tol
    e = tc.templates.measurements.pauli_sum_expectation(c, [ps for _, ps in terms], [w for w, _ in terms])
    np.testing.assert_allclose(e, sum(w * o.expectation_ps(ps=ps).real for w, ps in terms), rtol=10 * tol, atol=10 * tol)
    tc.set_dtype("complex64")


def test_random_circuit_small(eng):
    """Config 4 recipe at n = 9: amplitudes + identical sample indices."""
    n = 9
    ops = orc.random_circuit(n, 6, seed=3)
    c = _run_gatelist(n, ops)
    o = orc.run_gatelist(n, ops)
    psi = A(c.state())
    assert np.linalg.norm(psi - o.state()) / np.linalg.norm(o.state()) < 1e-5
    u = np.random.default_rng(4).random(2000)
    got = c.sample(batch=2000, allow_state=True, format="sample_int", status=u)
    cdf = orc.sample_cdf(np.abs(psi.astype(np.complex128)) ** 2)
    want = np.searchsorted(cdf, cdf[-1] * (1 - u), side="left")
    bad = np.nonzero(got != want)[0]
    r = cdf[-1] * (1 - u)
    for i in bad:  # only CDF ties within tolerance may differ
        assert abs(cdf[min(got[i], want[i])] - r[i]) < 1e-6, i


def test_incremental_execution(eng):
    """Gates added after a query are applied to the cached state (basecircuit.py:245 semantics)."""
    c = tc.Circuit(4)
    c.h(0)
    c.cnot(0, 1)
    s1 = A(c.state())
    c.rx(2, theta=0.3)
    c.cz(1, 2)
    o = OracleCircuit(4)
    o.h(0)
    o.cnot(0, 1)
    np.testing.assert_allclose(s1, o.state(), atol=1e-6)
    o.rx(2, theta=0.3)
    o.cz(1, 2)
    np.testing.assert_allclose(c.state(), o.state(), atol=1e-6)


# ---- vmap ---------------------------------------------------------------------------------------
def test_vmap_basic(eng):
    # tests/test_backends.py:23-52 (shape semantics) + value parity against a python loop
    K = tc.backend

    def f(theta):
        c = tc.Circuit(3)
        c.rx(0, theta=theta[0])
        c.ry(1, theta=theta[1])
        c.cnot(0, 2)
        c.rzz(1, 2, theta=theta[2] * 0.5)
        return K.real(c.expectation_ps(z=[2]) + 2.0 * c.expectation_ps(x=[1], z=[0]))

    th = np.random.default_rng(0).uniform(0, 2, size=(5, 3))
    got = K.vmap(f, vectorized_argnums=0)(th)
    assert got.shape == (5,)
    want = np.array([f(t) for t in th])
    np.testing.assert_allclose(got, want, atol=1e-5)


def test_vmap_hea_energy(eng):
    """Config 3 shape at n = 6, batch 7: vmapped TFIM energy == per-element oracle."""
    n, B = 6, 7
    params = np.random.default_rng(2).uniform(0, 2 * np.pi, size=[B, 3, 2, n])
    terms = orc.tfim_terms(n)
    pss = [ps for _, ps in terms]
    ws = [w for w, _ in terms]

    def energy(p):
        c = tc.Circuit(n)
        for l in range(3):
            for i in range(n):
                c.rx(i, theta=p[l, 0, i])
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=p[l, 1, i])
            for i in range(n - 1):
                c.cnot(i, i + 1)
        return tc.templates.measurements.pauli_sum_expectation(c, pss, ws)

    got = tc.backend.vmap(energy)(params)
    assert got.shape == (B,)
    for b in range(B):
        o = orc.run_gatelist(n, orc.hea_circuit(n, params[b]))
        want = sum(w * o.expectation_ps(ps=ps).real for w, ps in terms)
        np.testing.assert_allclose(got[b], want, atol=2e-5)
    # jit is the identity on this backend
    assert tc.backend.jit(energy) is energy


def test_vmap_batched_angle_types(eng):
    """Batched rotation / exp1 matrices come from the raw batch vector (gates._cos_sin_batched): integer,
    float32 and complex-typed angles, the half-angle of rzz / rxx and the full angle of exp1 all give the
    unbatched matrices; a vector angle per batch element is rejected (gates.py:463-636, 826-865)."""
    from tensorcircuit_b200 import gates
    from tensorcircuit_b200.batching import BatchArray

    th = np.array([0, 1, 2, 5])
    zz = np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0]))
    for conv in (lambda a: a, lambda a: a.astype(np.float32) * 0.37, lambda a: a * 0.21 + 0.0j, lambda a: a * (0.3 + 0.1j)):
        a = conv(th)
        for build in (lambda t: gates.rx_gate(theta=t), lambda t: gates.ry_gate(theta=t), lambda t: gates.rz_gate(theta=t),
                      lambda t: gates.rzz_gate(theta=t), lambda t: gates.rxx_gate(theta=t),
                      lambda t: gates.exp1_gate(unitary=zz, theta=t), lambda t: gates.exp1_gate(unitary=zz, theta=t, half=True)):
            got = build(BatchArray(a)).matrix().a
            assert got.shape[0] == len(th)
            for b in range(len(th)):
                t = a[b].item()
                if isinstance(t, complex) and t.imag == 0:
                    t = t.real
                np.testing.assert_allclose(np.asarray(got[b]), build(t).matrix(), atol=1e-13)
    with pytest.raises(ValueError):
        gates.rx_gate(theta=BatchArray(np.zeros((4, 2))))
    with pytest.raises(ValueError):
        gates.rzz_gate(theta=BatchArray(np.zeros((4, 3))))


# ---- pass planner --------------------------------------------------------------------------------
@pytest.mark.parametrize("rt_gates", [2, 12])
@pytest.mark.parametrize("dtype,tol", [("complex64", 1e-5), ("complex128", 1e-11)])
def test_planned_passes_production_tile(eng, dtype, tol, rt_gates, monkeypatch):
    """Default 64 KiB pass tile, 256 threads, several tiles: the unrolled / Gray-code paths of
    the pass kernels (register tiles in pair mode and in generic mode) against the oracle."""
    st_cls = tc.engine.DeviceState
    monkeypatch.setattr(st_cls, "regtile_max_gates", rt_gates)
    monkeypatch.setattr(st_cls, "use_regtiles", True, raising=False)
    tc.set_dtype(dtype)
    n = 15
    ops = orc.random_circuit(n, 3, seed=9)
    ops += [("toffoli", (0, 7, 14), {}), ("rzz", (3, 11), {"theta": 0.4}), ("h", (14,), {}), ("cz", (13, 14), {})]
    c = _run_gatelist(n, ops)
    o = orc.run_gatelist(n, ops)
    assert np.linalg.norm(A(c.state()) - o.state()) / np.linalg.norm(o.state()) < tol
    tc.set_dtype("complex64")



@pytest.mark.parametrize("kf", [2, 3, 4])
def test_planned_passes_small_tiles(eng, kf, monkeypatch):
    """Multi-block staged passes with a tiny tile (many passes, gathered high bits) == oracle."""
    from tensorcircuit_b200.fusion import fuse, plan_passes

    monkeypatch.setenv("TCB200_PASS_TILE_BYTES_LOG2", "9")  # 64 complex64 amplitudes per tile
    monkeypatch.setattr(tc.Circuit, "fusion_kmax", kf)
    monkeypatch.setattr(tc.Circuit, "use_passes", True)
    n = 11
    ops = orc.random_circuit(n, 4, seed=7) + orc.hea_circuit(n, np.random.default_rng(3).uniform(0, 6, size=[1, 2, n]))
    c = _run_gatelist(n, ops)
    o = orc.run_gatelist(n, ops)
    assert np.linalg.norm(A(c.state()) - o.state()) / np.linalg.norm(o.state()) < 2e-5
    # plan properties: every block exactly once, dependency order respected, geometry limits
    blocks = fuse(c._ops, n, kmax=kf)
    passes = plan_passes([b.bits for b in blocks], n, 6, max_hi=3)
    order = [i for p in passes for i in p.block_ids]
    assert sorted(order) == list(range(len(blocks)))
    pos = {b: i for i, b in enumerate(order)}
    last = {}
    for i, b in enumerate(blocks):
        for q in b.bits:
            if q in last:
                assert pos[last[q]] < pos[i]
            last[q] = i
    for p in passes:
        if len(p.block_ids) == 1:  # standalone blocks go through the single-block kernel
            continue
        assert len(p.tile_hi) <= 2 and len(p.block_ids) <= 16
        lrow = 6 - len(p.tile_hi)
        for i in p.block_ids:
            if len(blocks[i].bits) <= 4:
                assert all(q < lrow or q in p.tile_hi for q in blocks[i].bits)
    assert len(passes) < len(blocks)


# ---- Monte-Carlo trajectories (tests/test_circuit.py:120-230, 393-401, 772-796, 1439-1446, 1596-1643)
def test_jittable_depolarizing(eng):
    # tests/test_circuit.py:120-230: five spellings of a depolarizing trajectory keep the norm
    n = 5
    K = tc.backend

    def pre():
        c = tc.Circuit(n)
        for i in range(n):
            c.H(i)
        for i in range(n):
            c.cnot(i, (i + 1) % n)
        return c

    def f1():
        c = pre()
        for i in range(n):
            c.unitary_kraus([tc.gates._x_matrix, tc.gates._y_matrix, tc.gates._z_matrix, tc.gates._i_matrix], i, prob=[0.2, 0.2, 0.2, 0.4])
        for i in range(n):
            c.cz(i, (i + 1) % n)
        return c.wavefunction()

    def f2():
        c = pre()
        for i in range(n):
            c.unitary_kraus(tc.channels.depolarizingchannel(0.2, 0.2, 0.2), i)
        return c.wavefunction()

    def f3():
        c = pre()
        for i in range(n):
            c.depolarizing(i, px=0.2, py=0.2, pz=0.2)
        return c.wavefunction()

    def f4():
        c = pre()
        for i in range(n):
            c.depolarizing2(i, px=0.2, py=0.2, pz=0.2)
        return c.wavefunction()

    def f5():
        c = pre()
        for i in range(n):
            c.unitary_kraus2(tc.channels.depolarizingchannel(0.2, 0.2, 0.2), i)
        return c.wavefunction()

    for f in [f1, f2, f3, f4, f5]:
        K.set_random_state(23)
        np.testing.assert_allclose(K.norm(K.jit(f)()), 1.0, atol=1e-4)
        np.testing.assert_allclose(K.norm(K.jit(f)()), 1.0, atol=1e-4)


def test_postselection(eng):
    # tests/test_circuit.py:393-401
    c = tc.Circuit(3)
    c.H(1)
    c.H(2)
    c.mid_measurement(1, 1)
    c.mid_measurement(2, 1)
    s = c.wavefunction()
    np.testing.assert_allclose(A(s[3]).real, 0.5, atol=1e-6)


def test_teleportation(eng):
    # tests/test_circuit.py:772-796
    tc.backend.set_random_state(42)
    for _ in range(6):
        c = tc.Circuit(2)
        c.H(0)
        r = c.cond_measurement(0)
        c.conditional_gate(r, [tc.gates.i(), tc.gates.x()], 1)
        e = c.expectation([tc.gates.z(), [1]])
        np.testing.assert_allclose(e, -1 if A(r) > 0.5 else 1, atol=1e-5)


def test_channel_auto_register(eng, highp):
    # tests/test_circuit.py:1439-1446
    c = tc.Circuit(2)
    c.H(0)
    c.reset(0, status=0.8)
    s = c.state()
    np.testing.assert_allclose(A(s[0]), 1.0, atol=1e-9)


def test_general_kraus_with_prob(eng):
    # tests/test_circuit.py:1596-1643
    c = tc.Circuit(2)
    c.h([0, 1])
    p = 0.5
    status = [0.3, 0.8]
    rs = []
    for i in range(2):
        ks = [np.sqrt(p) * np.array([[1.0, 0], [0, 0]]), np.sqrt(p) * np.array([[0, 0], [0, 1.0]]), np.sqrt(1 - p) * np.eye(2)]
        rs.append(c.general_kraus(ks, i, status=status[i], with_prob=True))
    np.testing.assert_allclose(rs[0][0], 1)
    np.testing.assert_allclose(rs[1][0], 2)
    np.testing.assert_allclose(c.expectation_ps(z=[0]), -1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(z=[1]), 0, atol=1e-5)
    np.testing.assert_allclose(A(rs[0][1]), [0.25, 0.25, 0.5], atol=1e-5)
    np.testing.assert_allclose(A(rs[1][1]), [0.25, 0.25, 0.5], atol=1e-5)
    np.testing.assert_allclose(tc.backend.norm(c.state()), 1, atol=1e-5)


def test_channel_identity(eng):
    # tests/test_channels.py:24-48: every registered channel is trace preserving
    ch = tc.channels
    for ks in [ch.depolarizingchannel(0.1, 0.2, 0.3), ch.amplitudedampingchannel(0.25, 0.3), ch.phasedampingchannel(0.6), ch.resetchannel(),
               ch.generaldepolarizingchannel(0.02, 2), ch.isotropicdepolarizingchannel(0.3, 2), ch.generaldepolarizingchannel([0.1, 0.2, 0.3], 1)]:
        ch.kraus_identity_check(ks)


def test_depolarizing_trajectory_average(eng, highp):
    # the trajectory average reproduces the channel: E_status[|psi><psi|] = sum_k K rho K^dag
    # (tests/test_channels.py:100-135 checks the same through the density-matrix simulator)
    n, nt = 3, 400
    K = tc.backend

    def f(status):
        c = tc.Circuit(n)
        c.h(0)
        c.cnot(0, 1)
        c.ry(2, theta=0.7)
        c.depolarizing(1, px=0.1, py=0.2, pz=0.3, status=status[0])
        c.amplitudedamping(2, gamma=0.4, p=1.0, status=status[1])
        return K.real(c.expectation_ps(z=[1])), K.real(c.expectation_ps(z=[2])), K.real(c.expectation_ps(x=[0], z=[1]))

    st = np.random.default_rng(0).random((nt, 2))
    z1, z2, xz = [np.asarray(v) for v in K.vmap(f)(st)]
    # exact channel values from the oracle's density matrix
    o = OracleCircuit(n)
    o.h(0)
    o.cnot(0, 1)
    o.ry(2, theta=0.7)
    rho = np.outer(o.state(), o.state().conj())

    def chan(rho, ks, q):
        out = np.zeros_like(rho)
        for k in ks:
            full = np.kron(np.kron(np.eye(2**q), k), np.eye(2 ** (n - q - 1)))
            out += full @ rho @ full.conj().T
        return out

    rho = chan(rho, orc.ch_depolarizing(0.1, 0.2, 0.3), 1)
    rho = chan(rho, orc.ch_amplitudedamping(0.4, 1.0), 2)
    ez1 = np.real(np.trace(rho @ orc.pauli_string_matrix([0, 3, 0])))
    ez2 = np.real(np.trace(rho @ orc.pauli_string_matrix([0, 0, 3])))
    exz = np.real(np.trace(rho @ orc.pauli_string_matrix([1, 3, 0])))
    assert abs(z1.mean() - ez1) < 4.0 / np.sqrt(nt)
    assert abs(z2.mean() - ez2) < 4.0 / np.sqrt(nt)
    assert abs(xz.mean() - exz) < 4.0 / np.sqrt(nt)


def test_heisenberg_measurements(eng):
    # templates/measurements.py:211-287 docstring example (energy 1 for |10000> on an open... ring of 5)
    g = tc.templates.graphs.Line1D(n=5)
    c = tc.Circuit(5)
    c.X(0)
    np.testing.assert_allclose(tc.templates.measurements.heisenberg_measurements(c, g), 1.0, atol=1e-5)
    # generic state and couplings against the oracle, term by term
    n = 6
    g = tc.templates.graphs.Line1D(n, edge_weight=[0.5, 1.0, -0.7, 0.3, 1.2, 0.9], pbc=True)
    ops = orc.hea_circuit(n, np.random.default_rng(5).uniform(0, 6, size=[2, 2, n]))
    c = _run_gatelist(n, ops)
    o = orc.run_gatelist(n, ops)
    want = 0.0
    for a, b in g.edges:
        w = g[a][b]["weight"]
        want += w * (0.8 * o.expectation_ps(z=[a, b]) + 0.6 * o.expectation_ps(y=[a, b]) - 0.4 * o.expectation_ps(x=[a, b])).real
    for i in range(n):
        want += (0.3 * o.expectation_ps(x=[i]) - 0.2 * o.expectation_ps(z=[i])).real
    got = tc.templates.measurements.heisenberg_measurements(c, g, hzz=0.8, hyy=0.6, hxx=-0.4, hx=0.3, hz=-0.2)
    np.testing.assert_allclose(got, want, atol=2e-5)


# ---- gradients (backends/jax_backend.py:668-776; tests/test_backends.py:453-476) --------------------
def _vqe_energy_oracle(params, n, nlayers):
    o = OracleCircuit(n)
    for i in range(n):
        o.h(i)
    for l in range(nlayers):
        for i in range(n - 1):
            o.rzz(i, i + 1, theta=params[2 * l, i])
        for i in range(n):
            o.rx(i, theta=params[2 * l + 1, i])
    e = 0.0
    for i in range(n):
        e += o.expectation_ps(x=[i]).real
    for i in range(n - 1):
        e += 0.7 * o.expectation_ps(z=[i, i + 1]).real
    return e


def test_value_and_grad_vqe(eng, highp):
    """K.value_and_grad of a VQE energy (rzz / rx layers, sum of <X_i> and <Z_i Z_i+1>) against
    central differences of the float64 oracle."""
    n, nlayers = 4, 2
    K = tc.backend

    def energy(params, scale):
        c = tc.Circuit(n)
        for i in range(n):
            c.h(i)
        for l in range(nlayers):
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=params[2 * l, i])
            for i in range(n):
                c.rx(i, theta=params[2 * l + 1, i])
        e = 0.0
        for i in range(n):
            e = e + c.expectation_ps(x=[i])
        zz = tc.templates.measurements.pauli_sum_expectation(c, [[3 if q in (i, i + 1) else 0 for q in range(n)] for i in range(n - 1)])
        return K.real(e) + scale * zz + 0.1 * K.sum(params**2)  # the last term: direct dependence

    params = np.random.default_rng(3).uniform(0, 2, size=(2 * nlayers, n))
    v, g = K.value_and_grad(energy)(params, 0.7)
    want_v = _vqe_energy_oracle(params, n, nlayers) + 0.1 * np.sum(params**2)
    np.testing.assert_allclose(v, want_v, atol=1e-9)
    want_g = np.zeros_like(params)
    h = 1e-6
    for idx in np.ndindex(params.shape):
        p1, p2 = params.copy(), params.copy()
        p1[idx] += h
        p2[idx] -= h
        want_g[idx] = (_vqe_energy_oracle(p1, n, nlayers) - _vqe_energy_oracle(p2, n, nlayers)) / (2 * h) + 0.2 * params[idx]
    np.testing.assert_allclose(g, want_g, atol=2e-7)
    # grad alone, and the second argument (enters only the host arithmetic)
    np.testing.assert_allclose(K.grad(energy)(params, 0.7), want_g, atol=2e-7)
    v2, (g0, g1) = K.value_and_grad(energy, argnums=(0, 1))(params, 0.7)
    o = OracleCircuit(n)
    np.testing.assert_allclose(g0, want_g, atol=2e-7)
    zz_val = (want_v - 0.1 * np.sum(params**2) - _vqe_energy_oracle_x_only(params, n, nlayers)) / 0.7
    np.testing.assert_allclose(g1, zz_val, atol=1e-7)


def _vqe_energy_oracle_x_only(params, n, nlayers):
    o = OracleCircuit(n)
    for i in range(n):
        o.h(i)
    for l in range(nlayers):
        for i in range(n - 1):
            o.rzz(i, i + 1, theta=params[2 * l, i])
        for i in range(n):
            o.rx(i, theta=params[2 * l + 1, i])
    return sum(o.expectation_ps(x=[i]).real for i in range(n))


def test_vvag_shared_and_vectorized_args(eng):
    """vvag: values per batch element, gradient of the sum -- per element for the vectorised
    argument, summed for the shared one (tests/test_backends.py:453-476 semantics)."""
    n = 3
    K = tc.backend

    def f(x, w):
        c = tc.Circuit(n)
        for i in range(n):
            c.ry(i, theta=x[i])
        c.cnot(0, 1)
        c.rzz(1, 2, theta=w[0])
        c.rx(0, theta=w[1] * x[0])  # a parameter product: both arguments move this gate
        return K.real(c.expectation_ps(z=[0, 2]) + 0.5 * c.expectation_ps(x=[1]))

    def oracle(x, w):
        o = OracleCircuit(n)
        for i in range(n):
            o.ry(i, theta=x[i])
        o.cnot(0, 1)
        o.rzz(1, 2, theta=w[0])
        o.rx(0, theta=w[1] * x[0])
        return (o.expectation_ps(z=[0, 2]) + 0.5 * o.expectation_ps(x=[1])).real

    xs = np.random.default_rng(1).uniform(0, 3, size=(3, n))
    w = np.array([0.4, 1.3])
    vals, (gx, gw) = K.vvag(f, argnums=(0, 1), vectorized_argnums=0)(xs, w)
    h = 1e-6
    for b in range(3):
        np.testing.assert_allclose(vals[b], oracle(xs[b], w), atol=2e-6)
        for i in range(n):
            e = np.zeros(n)
            e[i] = h
            np.testing.assert_allclose(gx[b, i], (oracle(xs[b] + e, w) - oracle(xs[b] - e, w)) / (2 * h), atol=2e-5)
    for i in range(2):
        e = np.zeros(2)
        e[i] = h
        want = sum((oracle(xs[b], w + e) - oracle(xs[b], w - e)) / (2 * h) for b in range(3))
        np.testing.assert_allclose(gw[i], want, atol=5e-5)


def test_ad(eng):
    # tests/test_circuit.py:296-317 (universal_ad): scalar parameter, grad == value_and_grad
    K = tc.backend

    def forward(theta):
        c = tc.Circuit(2)
        c.R(0, theta=theta, alpha=0.5, phi=0.8)
        return K.real(c.expectation((tc.gates.z(), [0])))

    gg = K.jit(K.grad(forward))
    vg = K.jit(K.value_and_grad(forward))
    theta = tc.gates.num_to_tensor(1.0)
    grad1 = gg(theta)
    v2, grad2 = vg(theta)
    assert grad1 == grad2

    def want(t):
        o = OracleCircuit(2)
        o.r(0, theta=t, alpha=0.5, phi=0.8)
        return o.expectation_ps(z=[0]).real

    np.testing.assert_allclose(v2, want(1.0), atol=2e-6)
    np.testing.assert_allclose(grad2, (want(1.0 + 1e-6) - want(1.0 - 1e-6)) / 2e-6, atol=2e-5)


def test_vqe_training_loop(eng):
    """A VQE script as users write it against the reference (benchmarks/scripts/vqe_tc.py shape):
    K.jit(K.value_and_grad(energy)) inside a gradient-descent loop lowers the TFIM energy."""
    n, nlayers = 4, 2
    K = tc.backend
    g = tc.templates.graphs.Line1D(n, pbc=False)

    def energy(params):
        c = tc.Circuit(n)
        for i in range(n):
            c.h(i)
        for l in range(nlayers):
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=params[2 * l, i])
            for i in range(n):
                c.rx(i, theta=params[2 * l + 1, i])
        return tc.templates.measurements.heisenberg_measurements(c, g, hzz=1.0, hxx=0.0, hyy=0.0, hx=-1.0)

    vg = K.jit(K.value_and_grad(energy))
    params = 0.1 * np.ones((2 * nlayers, n))
    e0, _ = vg(params)
    for _ in range(15):
        e, grad = vg(params)
        params = params - 0.1 * grad
    e1, _ = vg(params)
    assert e1 < e0 - 0.5, (e0, e1)
    # exact ground energy of the open 4-site TFIM (J = 1, h = -1) bounds it from below
    H = np.zeros((16, 16), dtype=complex)
    for i in range(n - 1):
        H += orc.pauli_string_matrix([3 if q in (i, i + 1) else 0 for q in range(n)])
    for i in range(n):
        H -= orc.pauli_string_matrix([1 if q == i else 0 for q in range(n)])
    assert e1 > np.linalg.eigvalsh(H)[0] - 1e-4


def test_backend_small_tensor_table(eng):
    # tests/test_backends.py (method table smoke checks): host glue the scripts call around circuits
    K = tc.backend
    np.testing.assert_allclose(K.softmax(np.array([1.0, 2.0, 3.0])).sum(), 1.0)
    np.testing.assert_allclose(K.relu(np.array([-1.0, 2.0])), [0, 2])
    np.testing.assert_allclose(K.sigmoid(np.array([0.0])), [0.5])
    np.testing.assert_allclose(K.acos(K.cos(np.array([0.3]))), [0.3])
    np.testing.assert_allclose(K.atan2(np.array([1.0]), np.array([1.0])), [np.pi / 4])
    np.testing.assert_allclose(K.tile(np.array([1, 2]), [2]), [1, 2, 1, 2])
    np.testing.assert_allclose(K.std(np.array([1.0, 3.0])), 1.0)
    assert K.argmin(np.array([3, 1, 2])) == 1
    assert K.cond(True, lambda: 1, lambda: 2) == 1 and K.switch(1, [lambda: 1, lambda: 2]) == 2
    assert K.scan(lambda c, x: c + x, np.arange(4), 0) == 6
    np.testing.assert_allclose(K.eigvalsh(np.diag([2.0, 1.0])), [1, 2])
    np.testing.assert_allclose(K.sqrtmh(np.diag([4.0, 9.0])), np.diag([2.0, 3.0]), atol=1e-12)
    leaves, td = K.tree_flatten({"a": [np.ones(2), 3.0], "b": (1,)})
    assert len(leaves) == 3
    back = K.tree_unflatten(td, leaves)
    assert back["b"] == (1,) and back["a"][1] == 3.0
    assert K.tree_map(lambda x, y: x + y, [1, (2, 3)], [10, (20, 30)]) == [11, (22, 33)]
    K.set_random_state(7)
    r = K.implicit_randc(4, shape=[100], p=np.array([0.0, 1.0, 0.0, 0.0]))
    assert np.all(r == 1)


def test_value_and_grad_limits(eng):
    """what the gradient does not cover fails loudly, not silently"""
    K = tc.backend

    def uses_state(p):
        c = tc.Circuit(2)
        c.rx(0, theta=p[0])
        return K.real(K.sum(np.asarray(c.wavefunction())))

    with pytest.raises(NotImplementedError):
        K.value_and_grad(uses_state)(np.array([0.3]))

    def uses_inputs(p):
        c = tc.Circuit(1, inputs=np.array([1.0, 0.0]))
        c.rx(0, theta=p[0])
        return K.real(c.expectation_ps(z=[0]))

    with pytest.raises(NotImplementedError):
        K.value_and_grad(uses_inputs)(np.array([0.3]))
    with pytest.raises(NotImplementedError):
        K.hessian(lambda x: x, argnums=(0, 1))  # (hessian / jacobians of one argument exist: tests/test_operators.py)
    with pytest.raises(ValueError):
        K.value_and_grad(lambda p: p * np.ones(2))(np.array([0.3]))  # not a scalar loss
    # a loss that does not touch a circuit at all still differentiates (direct dependence only)
    v, g = K.value_and_grad(lambda p: K.sum(p**2))(np.array([1.0, 2.0]))
    np.testing.assert_allclose(v, 5.0)
    np.testing.assert_allclose(g, [2.0, 4.0], atol=1e-7)


def test_circuit_copy_prepend_instructions(eng):
    # tests/test_circuit.py:1650-1655 (copy), abstractcircuit.py:1133-1146 (prepend), :655-746
    c = tc.Circuit(2)
    c.h(0)
    c1 = c.copy()
    c.rz(0, theta=0.1)
    assert c1.gate_count() == 1 and c.gate_count() == 2
    c2 = tc.Circuit(2)
    c2.x(1)
    c.prepend(c2)  # X(1) first, then H(0), rz(0)
    assert [d["name"] for d in c.to_qir()] == ["x", "h", "rz"]
    o = OracleCircuit(2)
    o.x(1)
    o.h(0)
    o.rz(0, theta=0.1)
    np.testing.assert_allclose(A(c.state()), o.state(), atol=1e-6)
    assert c.gate_count_by_condition(lambda d: d["index"] == (0,)) == 2
    c.measure_instruction(0, 1)
    c.barrier_instruction(0, 1)
    c.reset_instruction(1)
    assert [d["name"] for d in c._extra_qir] == ["measure", "measure", "barrier", "reset"]
    np.testing.assert_allclose(A(c.state()), o.state(), atol=1e-6)  # instructions do not touch the state
    assert c.is_valid()
    c.select_gate(1, [tc.gates.i(), tc.gates.x()], 0)
    o.x(0)
    np.testing.assert_allclose(A(c.state()), o.state(), atol=1e-6)


def test_value_and_grad_has_aux_and_shared_parameter(eng):
    """has_aux passes the auxiliary output through; one parameter feeding several gates (and a
    non-linear loss) gets the sum of its gate derivatives."""
    K = tc.backend

    def f(t):
        c = tc.Circuit(3)
        for i in range(3):
            c.ry(i, theta=t[0] * (i + 1))   # t[0] moves three gates
        c.cnot(0, 1)
        c.rzz(1, 2, theta=t[1] ** 2)         # non-linear use of t[1]
        e = K.real(c.expectation_ps(z=[2]))
        return (e - 0.3) ** 2, {"energy": e}  # non-linear host arithmetic + aux

    def want(t):
        o = OracleCircuit(3)
        for i in range(3):
            o.ry(i, theta=t[0] * (i + 1))
        o.cnot(0, 1)
        o.rzz(1, 2, theta=t[1] ** 2)
        return (o.expectation_ps(z=[2]).real - 0.3) ** 2

    t = np.array([0.4, 0.9])
    (v, aux), g = K.value_and_grad(f, has_aux=True)(t)
    np.testing.assert_allclose(v, want(t), atol=2e-6)
    np.testing.assert_allclose(np.sqrt(v), abs(float(aux["energy"]) - 0.3), atol=2e-6)
    h = 1e-6
    fd = [(want(t + h * np.eye(2)[k]) - want(t - h * np.eye(2)[k])) / (2 * h) for k in range(2)]
    np.testing.assert_allclose(g, fd, atol=3e-5)


def test_value_and_grad_sample_expectation_two_bases(eng, highp):
    """queries on the same circuit in different measurement bases must not share a shifted
    simulation (advisor finding: the group key was (id(circ), len(ops)))"""
    import gc

    K = tc.backend

    def f(t):
        c = tc.Circuit(3)
        for i in range(3):
            c.ry(i, theta=t[i])
            c.rz(i, theta=t[3 + i])
        c.cnot(0, 1)
        c.cnot(1, 2)
        e = c.sample_expectation_ps(x=[0]) + 2.0 * c.sample_expectation_ps(y=[0]) + 0.5 * c.sample_expectation_ps(x=[1])
        gc.collect()
        return e

    def want(t):
        o = OracleCircuit(3)
        for i in range(3):
            o.ry(i, theta=t[i])
            o.rz(i, theta=t[3 + i])
        o.cnot(0, 1)
        o.cnot(1, 2)
        return (o.expectation_ps(x=[0]) + 2.0 * o.expectation_ps(y=[0]) + 0.5 * o.expectation_ps(x=[1])).real

    t = np.array([0.3, 0.7, 1.1, 0.5, 0.2, 0.9])
    v, g = K.value_and_grad(f)(t)
    np.testing.assert_allclose(v, want(t), atol=1e-9)
    h = 1e-6
    fd = [(want(t + h * np.eye(6)[k]) - want(t - h * np.eye(6)[k])) / (2 * h) for k in range(6)]
    np.testing.assert_allclose(g, fd, atol=2e-6)


def test_value_and_grad_subcircuits_in_a_loop(eng, highp):
    """several circuits built (and dropped) inside the loss: object ids may be reused"""
    import gc

    K = tc.backend

    def f(t):
        e = 0.0
        for k in range(3):
            c = tc.Circuit(2)
            c.rx(0, theta=t[k])
            c.cnot(0, 1)
            c.ry(1, theta=t[(k + 1) % 3])
            e = e + (k + 1) * K.real(c.expectation_ps(z=[1]))
            del c
            gc.collect()
        return e

    def want(t):
        e = 0.0
        for k in range(3):
            o = OracleCircuit(2)
            o.rx(0, theta=t[k])
            o.cnot(0, 1)
            o.ry(1, theta=t[(k + 1) % 3])
            e += (k + 1) * o.expectation_ps(z=[1]).real
        return e

    t = np.array([0.4, 1.3, 0.8])
    v, g = K.value_and_grad(f)(t)
    np.testing.assert_allclose(v, want(t), atol=1e-9)
    h = 1e-6
    fd = [(want(t + h * np.eye(3)[k]) - want(t - h * np.eye(3)[k])) / (2 * h) for k in range(3)]
    np.testing.assert_allclose(g, fd, atol=2e-6)


def test_qaoa_block_per_edge_parameters(eng):
    """templates/blocks.py:84-110: vector paramzz / paramx index edges / nodes"""
    import networkx as nx

    g = nx.Graph()
    g.add_edge(0, 1, weight=1.0)
    g.add_edge(1, 2, weight=2.0)
    pz, px = np.array([0.3, 0.8]), np.array([0.2, 0.5, 0.9])
    c = tc.Circuit(3)
    for i in range(3):
        c.h(i)
    tc.templates.blocks.QAOA_block(c, g, pz, px)
    o = OracleCircuit(3)
    for i in range(3):
        o.h(i)
    zz = np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0]))
    for i, (a, b) in enumerate(g.edges):
        o.exp1(a, b, unitary=zz, theta=pz[i])
    for i, nd in enumerate(g.nodes):
        o.rx(nd, theta=px[i])
    np.testing.assert_allclose(A(c.state()), o.state(), atol=1e-6)
    # a shared scalar angle is scaled by the edge weight
    c2 = tc.Circuit(3)
    tc.templates.blocks.QAOA_block(c2, g, 0.4, 0.7)
    o2 = OracleCircuit(3)
    for a, b in g.edges:
        o2.exp1(a, b, unitary=zz, theta=0.4 * g[a][b]["weight"])
    for nd in g.nodes:
        o2.rx(nd, theta=0.7)
    np.testing.assert_allclose(A(c2.state()), o2.state(), atol=1e-6)
    with pytest.raises(ValueError):
        tc.Circuit(2).rx(0, theta=np.array([0.1, 0.2]))


def test_sample_expectation_ps_bad_index_leaves_record_clean(eng):
    c = tc.Circuit(3)
    c.h(0)
    c.cnot(0, 1)
    n_ops, n_qir = len(c._ops), len(c._qir)
    with pytest.raises(ValueError):
        c.sample_expectation_ps(x=[0, 7], shots=16, status=np.random.default_rng(0).random(16))
    assert len(c._ops) == n_ops and len(c._qir) == n_qir
    np.testing.assert_allclose(c.sample_expectation_ps(x=[0, 1]), 1.0, atol=1e-6)


def test_expectation_ps_loop_is_coalesced(eng):
    """The reference idiom -- a Python loop over c.expectation_ps (examples/vqe_parallel_pmap.py:
    28-34) -- evaluates all its terms in ONE group when the energy is finally read (lazy.py)."""
    K = tc.backend
    n = 8
    rng = np.random.default_rng(3)
    th = rng.uniform(0, 2 * np.pi, size=(2, n))

    def build(cls):
        c = cls(n)
        for i in range(n):
            c.ry(i, theta=th[0, i])
        for i in range(n - 1):
            c.cnot(i, i + 1)
        for i in range(n):
            c.rx(i, theta=th[1, i])
        return c

    c = build(tc.Circuit)
    st = c._ensure_state()
    calls = []
    orig = st.expectation_terms
    st.expectation_terms = lambda fl, sg, ny: (calls.append(len(fl)), orig(fl, sg, ny))[1]
    e = 0.0
    for i in range(n):
        e += -1.0 * c.expectation_ps(x=[i])
    for i in range(n - 1):
        e = e + c.expectation_ps(z=[i, i + 1]) * 0.5
    e = K.real(e) / 2
    assert calls == []  # nothing evaluated yet
    o = build(OracleCircuit)
    want = (sum(-o.expectation_ps(x=[i]).real for i in range(n)) + 0.5 * sum(o.expectation_ps(z=[i, i + 1]).real for i in range(n - 1))) / 2
    np.testing.assert_allclose(float(e), want, atol=2e-5)
    assert calls == [2 * n - 1]  # one group for all terms
    # a value read before further gates stays valid; the new gate flushes what is pending
    a = c.expectation_ps(z=[0])
    c.x(0)
    b = c.expectation_ps(z=[0])
    np.testing.assert_allclose(float(np.real(a)), o.expectation_ps(z=[0]).real, atol=2e-5)
    np.testing.assert_allclose(float(np.real(b)), -o.expectation_ps(z=[0]).real, atol=2e-5)
    # eager mode agrees term by term
    c2 = build(tc.Circuit)
    c2.lazy_expectation = False
    v = c2.expectation_ps(x=[1], z=[2])
    assert isinstance(v, np.ndarray)
    np.testing.assert_allclose(v, o.expectation_ps(x=[1], z=[2]), atol=2e-5)
    np.testing.assert_allclose(build(tc.Circuit).expectation_ps(x=[1], z=[2]), v, atol=1e-7)


def test_expectation_loop_is_coalesced(eng):
    """The benchmark scripts' idiom -- a loop over c.expectation((gates.x(), [i])) and two-operator calls
    (benchmarks/scripts/vqe_tc.py:74-81, ``tfi_energy``) -- joins the same pool of pending terms as expectation_ps:
    one launch group when the energy is read."""
    K = tc.backend
    n = 7
    th = np.random.default_rng(5).uniform(0, 2 * np.pi, size=(2, n))

    def build(cls):
        c = cls(n)
        for i in range(n):
            c.ry(i, theta=th[0, i])
        for i in range(n - 1):
            c.cnot(i, i + 1)
        for i in range(n):
            c.rx(i, theta=th[1, i])
        return c

    def tfi_energy(c, j=1.0, h=-1.0):  # body of vqe_tc.py:74-81
        e = 0.0
        for i in range(n):
            e += h * c.expectation((tc.gates.x(), [i]))
        for i in range(n - 1):
            e += j * c.expectation((tc.gates.z(), [i]), (tc.gates.z(), [(i + 1) % n]))
        return e

    c = build(tc.Circuit)
    st = c._ensure_state()
    calls = []
    orig = st.expectation_terms
    st.expectation_terms = lambda fl, sg, ny: (calls.append(len(fl)), orig(fl, sg, ny))[1]
    e = K.real(tfi_energy(c))
    assert calls == []
    o = build(OracleCircuit)
    want = sum(-o.expectation_ps(x=[i]).real for i in range(n)) + sum(o.expectation_ps(z=[i, i + 1]).real for i in range(n - 1))
    np.testing.assert_allclose(float(e), want, atol=2e-5)
    assert calls == [2 * n - 1]
    # a non-Pauli operator (a projector: two Pauli terms) goes through the same pool
    p0 = tc.gates.Gate(np.array([[1.0, 0.0], [0.0, 0.0]]))
    np.testing.assert_allclose(float(np.real(c.expectation((p0, [2])))), 0.5 * (1 + o.expectation_ps(z=[2]).real), atol=2e-5)


def test_set_distributed_sample_order():
    """tc.set_distributed(..., sample_order=): "logical" = single-GPU indices for the same uniforms (dist.py restore_identity),
    "physical" = in place; anything else is an error.  (The sharded run itself: tests/test_dist_gloo.py, tests/test_dist_gpu.py.)"""
    from tensorcircuit_b200.dist import DistState

    old = DistState.sample_order
    try:
        tc.set_distributed(False, sample_order="logical")
        assert DistState.sample_order == "logical"
        tc.set_distributed(False, sample_order="physical")
        assert DistState.sample_order == "physical"
        with pytest.raises(ValueError):
            tc.set_distributed(False, sample_order="sorted")
    finally:
        DistState.sample_order = old
        tc.set_distributed(False)
