"""Pin the CPU oracle against the golden values held by the reference's own test-suite.

Every test names the reference test (file:line under /root/reference/) whose literal values it
transcribes.  These run on CPU (``-m "not gpu"``)."""

import numpy as np
import pytest

from oracle import tc_oracle as orc
from oracle.tc_oracle import OracleCircuit


def test_index_convention():
    # tests/test_circuit.py:22-43
    u = np.arange(16).reshape(2, 2, 2, 2)
    c = OracleCircuit(2)
    c.unitary(0, 1, unitary=u)
    assert c.state()[2].real == 8
    c = OracleCircuit(2)
    c.unitary(1, 0, unitary=u)
    assert c.state()[2].real == 4
    c = OracleCircuit(2)
    c.unitary(0, unitary=np.arange(4).reshape(2, 2))
    assert c.state()[2].real == 2


def test_basic_amplitudes():
    # tests/test_circuit.py:47-52
    c = OracleCircuit(2)
    c.x(0)
    assert c.state()[0b10] == 1.0
    c.CNOT(0, 1)
    assert c.state()[0b11] == 1.0
    # tests/test_circuit.py:80-84
    c = OracleCircuit(1)
    c.X(0)
    c.SD(0)
    np.testing.assert_allclose(c.state(), np.array([0.0, -1.0j]))
    # tests/test_circuit.py:319-323
    c = OracleCircuit(1)
    c.H(0)
    np.testing.assert_allclose(c.state(), np.array([1, 1]) / np.sqrt(2), atol=1e-4)


def test_inputs_unitary_form():
    # tests/test_circuit.py:404-412 : 2^(2n) inputs, extra trailing legs
    c = OracleCircuit(2, inputs=np.eye(4))
    c.X(0)
    c.Y(1)
    np.testing.assert_allclose(c.state().reshape(4, 4), np.kron(orc.X, orc.Y), atol=1e-4)
    # tests/test_circuit.py:64-68
    c = OracleCircuit(2, inputs=np.eye(4))
    c.iswap(0, 1)
    ans = np.array([[1.0, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1.0]])
    np.testing.assert_allclose(c.state().reshape(4, 4), ans, atol=1e-5)


def _matrix_of(n, build):
    c = OracleCircuit(n, inputs=np.eye(2**n))
    build(c)
    return c.state().reshape(2**n, 2**n)


def test_gate_values():
    # tests/test_gates.py:18-22
    c = OracleCircuit(1)
    c.h(0)
    c.phase(0, theta=np.pi / 2)
    np.testing.assert_allclose(c.state()[1], 0.7071j, atol=1e-4)
    # tests/test_gates.py:25-31
    m = _matrix_of(2, lambda c: c.cu(0, 1, theta=np.pi / 2, phi=-np.pi / 4, lbd=np.pi / 4))
    np.testing.assert_allclose(m[2:, 2:], orc.WROOT, atol=1e-5)
    np.testing.assert_allclose(m[:2, :2], np.eye(2), atol=1e-5)
    # tests/test_gates.py:60-77 (fsim numbers)
    def fsim(c):
        c.iswap(0, 1, theta=-0.2)
        c.cphase(0, 1, theta=-0.3)

    ans = np.array(
        [
            [1.0, 0.0, 0.0, 0.0],
            [0.0, 0.95105654, -0.309017j, 0.0],
            [0.0, -0.309017j, 0.95105654, 0.0],
            [0.0, 0.0, 0.0, 0.9553365 - 0.29552022j],
        ]
    )
    np.testing.assert_allclose(_matrix_of(2, fsim), ans, atol=1e-5)
    # tests/test_gates.py:80-91
    c = OracleCircuit(2)
    c.exp(0, 1, unitary=np.diag([1.0, -1, -1, 1]), theta=np.pi / 2)
    np.testing.assert_allclose(c.state()[0], -1j, atol=1e-12)
    # tests/test_gates.py:100-105
    np.testing.assert_allclose(orc.m_iswap(), np.array([[1.0, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1.0]]), atol=1e-5)
    np.testing.assert_allclose(orc.m_iswap(theta=0), np.eye(4), atol=1e-5)
    # tests/test_gates.py:137-141
    np.testing.assert_allclose(orc.FIXED["sd"], orc.S.conj().T)
    # tests/test_gates.py:13-16 : r gate == expm form
    th, al, ph = 1, 2, 3
    gen = np.sin(al) * np.cos(ph) * orc.X + np.sin(al) * np.sin(ph) * orc.Y + np.cos(al) * orc.Z
    np.testing.assert_allclose(orc.m_r(th, al, ph), orc.m_exp(gen, th), atol=1e-12)


def test_controlled_gates():
    # tests/test_gates.py:108-122
    np.testing.assert_allclose(orc.controlled(orc.controlled(orc.X)), orc.TOFFOLI)
    ocx = orc.ocontrolled(orc.controlled(orc.X))
    c = OracleCircuit(3)
    c.x(0)
    c.any(1, 0, 2, unitary=ocx)
    np.testing.assert_allclose(c.expectation((orc.Z, [2])), -1, atol=1e-5)
    # tests/test_gates.py:125-134 and tests/test_circuit.py:71-77
    c = OracleCircuit(2)
    c.x(0)
    c.crx(0, 1, theta=0.3)
    np.testing.assert_allclose(c.expectation((orc.Z, [1])), 0.95533645, atol=1e-5)
    c = OracleCircuit(2)
    c.x(1)
    c.crx(1, 0, theta=0.3)
    np.testing.assert_allclose(c.expectation((orc.Z, 0)), 0.95533645, atol=1e-5)


def test_rxx_family():
    # tests/test_gates.py:144-155
    c1 = OracleCircuit(3)
    c1.rxx(0, 1, theta=1.0)
    c1.ryy(0, 2, theta=0.5)
    c1.rzz(0, 1, theta=-0.5)
    c2 = OracleCircuit(3)
    c2.exp1(0, 1, theta=1.0 / 2, unitary=np.kron(orc.X, orc.X))
    c2.exp1(0, 2, theta=0.5 / 2, unitary=np.kron(orc.Y, orc.Y))
    c2.exp1(0, 1, theta=-0.5 / 2, unitary=np.kron(orc.Z, orc.Z))
    np.testing.assert_allclose(c1.state(), c2.state(), atol=1e-5)


def test_expectations():
    # tests/test_circuit.py:240-245
    c = OracleCircuit(2)
    c.H(0)
    np.testing.assert_allclose(c.expectation((orc.Z, [0])), 0, atol=1e-7)
    # tests/test_circuit.py:416-429
    c = OracleCircuit(2)
    c.X(0)
    np.testing.assert_allclose(c.expectation_ps(z=[0, 1]), -1, atol=1e-5)
    c = OracleCircuit(2)
    c.H(0)
    np.testing.assert_allclose(c.expectation_ps(z=[1], x=[0]), 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(ps=[1, 3]), 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(z=[1, 2], ps=[1, 3]), 1, atol=1e-5)
    # tests/test_circuit.py:1278-1281 (sign of Y)
    c = OracleCircuit(1, inputs=1 / np.sqrt(2) * np.array([-1, 1.0j]))
    np.testing.assert_allclose(c.expectation_ps(y=[0]), -1, atol=1e-5)
    # tests/test_circuit.py:1373-1380 (negative indices)
    c = OracleCircuit(3)
    c.H(-2)
    c.H(0)
    np.testing.assert_allclose(c.expectation_ps(x=[0]).real, 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(x=[1]).real, 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(x=[-1]).real, 0, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(z=[-2]).real, 0, atol=1e-5)
    # duplicate site -> ValueError (basecircuit.py:306-307)
    with pytest.raises(ValueError):
        c.expectation_ps(x=[0], z=[0])


def test_qir_append_value():
    # tests/test_circuit.py:707-727 : the three gates applied twice -> <Z1> = 0.202728
    c = OracleCircuit(3)
    for _ in range(2):
        c.H(0)
        c.rx(1, theta=0.7)
        c.exp1(0, 1, unitary=np.kron(orc.Z, orc.Z), theta=-0.2)
    np.testing.assert_allclose(c.expectation((orc.Z, [1])), 0.202728, atol=1e-5)


def test_mixed_measurement_four_vector():
    # tests/test_circuit.py:500-554 : <X_i> for i = 0..3
    n = 4
    c = OracleCircuit(n)
    for i in range(n):
        c.H(i)
    for _ in range(2):
        for i in range(n):
            c.cnot(i, (i + 1) % n)
        for i in range(n):
            c.rz(i, theta=1.0)
    v = [c.expectation_ps(x=[i]).real for i in range(n)]
    np.testing.assert_allclose(v, [0.157729, 0.157729, 0.157728, 0.085221], atol=1e-5)


def test_replace_inputs_value():
    # tests/test_circuit.py:571-580
    n = 3
    even = np.ones(2**n) / np.sqrt(2**n)
    c = OracleCircuit(n, inputs=even)
    for i in range(n):
        c.H(i)
    for i in range(n):
        np.testing.assert_allclose(c.expectation_ps(z=[i]), 1.0, atol=1e-5)


def test_inverse_roundtrip():
    # tests/test_circuit.py:1314-1343 : circuit followed by its inverse is the identity
    rng = np.random.default_rng(0)
    inputs = rng.uniform(size=8)
    inputs /= np.linalg.norm(inputs)
    c = OracleCircuit(3, inputs=inputs)
    c.iswap(0, 1)
    c.iswap(1, 0, theta=0.6)
    c.rxx(1, 2, theta=-0.2)
    c.cu(0, 1, lbd=2.0, theta=-0.7)
    c.r(2, alpha=0.3)
    c.sd(2)
    c.cx(1, 2)
    c.unitary(0, unitary=orc.X)
    for name, q, p in reversed(list(c.ops)):
        u = orc.gate_matrix(name, **p)
        c.any(*q, unitary=u.conj().T)
    np.testing.assert_allclose(c.state(), inputs, atol=1e-5)


def test_bell_block():
    # tests/test_templates.py:111-119 (templates/blocks.py:46-68)
    c = OracleCircuit(2, inputs=np.array([1.0, 0, 0, 0]))
    c.X(0)
    c.H(0)
    c.cnot(0, 1)
    c.X(1)
    np.testing.assert_allclose(c.state(), np.array([0.0, 0.70710677, -0.70710677, 0]), atol=1e-5)


def test_parameterized_measurement_values():
    # tests/test_templates.py:17-26
    c = OracleCircuit(2)
    c.H(0)
    c.H(1)
    np.testing.assert_allclose(c.expectation_ps(ps=[1, 1]), 1.0, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(ps=[3, 0]), 0.0, atol=1e-5)
    # tests/test_templates.py:29-39
    c = OracleCircuit(3)
    c.X(0)
    c.cnot(0, 1)
    c.H(-1)
    r = [c.expectation_ps(ps=[3 if j == i else 0 for j in range(3)][:2] + ([1] if i == 2 else [0])) for i in range(3)]
    np.testing.assert_allclose(np.real(r), [-1, -1, 1], atol=1e-5)


def test_pauli_string_matrices():
    # tests/test_miscs.py:26-55 : pins the closed form of quantum.py:1461-1482
    i, x, y, z = orc.PAULI
    pairs = [
        ([0, 0], np.eye(4)),
        ([0, 1], np.kron(i, x)),
        ([2, 1], np.kron(y, x)),
        ([3, 1], np.kron(z, x)),
        ([3, 2, 2, 0], np.kron(np.kron(np.kron(z, y), y), i)),
        ([0, 1, 1, 1], np.kron(np.kron(np.kron(i, x), x), x)),
    ]
    for ps, a in pairs:
        np.testing.assert_allclose(orc.pauli_string_matrix(ps), a, atol=1e-5)
    s = orc.pauli_string_matrix(pairs[4][0], 0.5) + orc.pauli_string_matrix(pairs[5][0], 1.0)
    np.testing.assert_allclose(s, 0.5 * pairs[4][1] + pairs[5][1], atol=1e-5)
    # closed form == operator application for random strings
    rng = np.random.default_rng(1)
    n = 5
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    for _ in range(10):
        ps = list(rng.integers(0, 4, size=n))
        m = orc.pauli_string_matrix(ps)
        x_, y_, z_ = orc.resolve_ps(n, ps=ps)
        np.testing.assert_allclose(orc.pauli_expectation(psi, n, x_, y_, z_), np.vdot(psi, m @ psi), atol=1e-10)


def test_sample_formats():
    # tests/test_quantum.py:299-312
    np.testing.assert_allclose(orc.spin_by_basis(2, 1), np.array([1, -1, 1, -1]))
    state = np.array([0.6, 0.4, 0, 0])
    np.testing.assert_allclose(orc.correlation_from_counts([0, 1], state), 0.2, atol=1e-5)
    np.testing.assert_allclose(orc.correlation_from_counts([1], state), 0.2, atol=1e-5)
    np.testing.assert_allclose(orc.correlation_from_samples([0, 1], np.array([0, 0, 3, 3, 3]), n=2), 1, atol=1e-5)
    # tests/test_quantum.py:456-464
    x, y = orc.count_d2s(np.array([0.1, 0, -0.3, 0]))
    np.testing.assert_allclose(x, [0, 2])
    np.testing.assert_allclose(y, [0.1, -0.3])
    np.testing.assert_allclose(orc.count_s2d((x, y), 2), [0.1, 0, -0.3, 0])
    # tests/test_quantum.py:467-489 (shapes of all formats)
    n = 4
    s = orc.probability_sample(np.ones(2**n), np.random.default_rng(0).random(9))
    assert orc.sample2all(s, n, "sample_bin").shape == (9, n)
    assert orc.sample2all(s, n, "sample_int").shape == (9,)
    assert orc.sample2all(s, n, "count_vector").shape == (2**n,)
    assert sum(orc.sample2all(s, n, "count_dict_bin").values()) == 9
    assert sum(orc.sample2all(s, n, "count_dict_int").values()) == 9
    np.testing.assert_array_equal(orc.sample_bin2int(orc.sample_int2bin(s, n), n), s)
    # MSB = qubit 0 (quantum.py:2104-2119)
    np.testing.assert_array_equal(orc.sample_int2bin(np.array([4]), 3), [[1, 0, 0]])


def test_sampler_rule():
    # tests/test_backends.py:278-280 : searchsorted side="left"
    np.testing.assert_array_equal(np.searchsorted([-1, 3.3, 9.1, 10.0], np.array([0.0, 4.1, 12.0], dtype=np.float32)), [1, 2, 4])
    # tests/test_backends.py:281-291 : statistical check of probability_sample
    p = np.array([0.05] * 8 + [0.2, 0.4])
    r = orc.probability_sample(p, np.random.default_rng(0).uniform(size=10000))
    _, cnt = np.unique(r, return_counts=True)
    np.testing.assert_allclose(cnt - p * 10000.0, np.zeros(10), atol=200)
    # (1 - u) and side="left" edge cases (abstract_backend.py:1145-1157)
    p = np.array([0.0, 0.25, 0.25, 0.5])
    assert orc.probability_sample(p, [0.0])[0] == 3  # r = total -> last index reaching it
    assert orc.probability_sample(p, [0.999999])[0] == 1  # r -> 0+, skips the leading zero
    assert orc.probability_sample(p, [0.5])[0] == 2  # r = 0.5 == cdf[2] -> left side
    assert orc.probability_sample(p, [0.75])[0] == 1


def test_sample_expectation_consistency():
    # tests/test_circuit.py:1383-1405 : sampled <ZZ> agrees with expectation_ps
    c = OracleCircuit(3)
    c.H(0)
    c.cnot(0, 1)
    c.rx(2, theta=0.4)
    s = c.sample_int(np.random.default_rng(3).random(200000))
    est = orc.correlation_from_samples([0, 1], s, 3)
    np.testing.assert_allclose(est, c.expectation_ps(z=[0, 1]).real, atol=1e-2)
    est = orc.correlation_from_samples([2], s, 3)
    np.testing.assert_allclose(est, c.expectation_ps(z=[2]).real, atol=1e-2)


def test_vmap_semantics_value():
    # tests/test_backends.py:23-52 analogue: stack per-element results on axis 0
    def f(theta):
        c = OracleCircuit(2)
        c.rx(0, theta=theta)
        return c.expectation_ps(z=[0]).real

    th = np.linspace(0, 1, 5)
    np.testing.assert_allclose([f(t) for t in th], np.cos(th), atol=1e-12)
