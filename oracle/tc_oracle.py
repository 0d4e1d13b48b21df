"""
CPU oracle for the TensorCircuit statevector hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a numpy (complex128) restatement of what the reference computes on the path
``tc.Circuit`` gate application -> ``wavefunction()`` -> ``expectation_ps()`` -> ``sample(status=)``.
It is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The engine in
``tensorcircuit_b200/`` never imports it and has no CPU path.

Pinning status:
  * tests/test_oracle_golden.py -- conventions (bit order, gate axes, Pauli signs, gate
    matrices, sample formats) against the golden values the reference's own test-suite holds
    (/root/reference/tests/test_circuit.py, test_gates.py, test_quantum.py, test_miscs.py,
    test_backends.py, test_templates.py -- each golden test cites the line it transcribes);
  * tests/test_ref_fixtures.py -- against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/
    ref_fixtures.npz is produced by oracle/make_golden.py, which runs the reference's own
    gates.py / abstractcircuit.py / basecircuit.py / circuit.py / quantum.py / backends from
    /root/reference unmodified (only the un-vendored tensornetwork / opt_einsum / graphviz
    imports are served by the stand-ins in oracle/refshim/).  Amplitudes (complex64 and
    complex128), expectation_ps, general-operator expectations, every gate name, sample
    formats, numpy-backend vmap -- and the sample indices for given ``status`` uniforms,
    which no reference test pins: the oracle reproduces them bit for bit (float32 CDF for
    complex64 states, float64 for complex128, abstract_backend.py:1124-1157).
The arithmetic of the reference lives in un-vendored third-party packages
(tensornetwork==0.4.6 per requirements/requirements-docker-v2.txt:7, opt_einsum, numpy/jax);
their published semantics (tensordot of the shared axes, transpose on reorder_edges) are what
``apply_gate`` below restates.

All citations are relative to /root/reference/.
"""

from __future__ import annotations

import math
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg

CDT = np.complex128

# ----------------------------------------------------------------------------------------
# fixed matrices  (tensorcircuit/gates.py:31-127)
# ----------------------------------------------------------------------------------------
_S2 = 1.0 / math.sqrt(2.0)
I2 = np.eye(2, dtype=CDT)
X = np.array([[0, 1], [1, 0]], dtype=CDT)
Y = np.array([[0, -1j], [1j, 0]], dtype=CDT)
Z = np.array([[1, 0], [0, -1]], dtype=CDT)
H = _S2 * np.array([[1, 1], [1, -1]], dtype=CDT)
S = np.diag([1, 1j]).astype(CDT)
T = np.diag([1, np.exp(0.25j * np.pi)]).astype(CDT)
# gates.py:39-43
WROOT = _S2 * np.array([[1, -_S2 * (1 + 1j)], [_S2 * (1 - 1j), 1]], dtype=CDT)
PAULI = [I2, X, Y, Z]


def _blockdiag(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros((a.shape[0] + b.shape[0],) * 2, dtype=CDT)
    out[: a.shape[0], : a.shape[0]] = a
    out[a.shape[0] :, a.shape[0] :] = b
    return out


def controlled(u: np.ndarray) -> np.ndarray:
    """[[I,0],[0,U]], new control is the first leg (gates.py:294-311)."""
    return _blockdiag(np.eye(u.shape[0], dtype=CDT), u)


def ocontrolled(u: np.ndarray) -> np.ndarray:
    """[[U,0],[0,I]], new control is the first leg (gates.py:313-331)."""
    return _blockdiag(u, np.eye(u.shape[0], dtype=CDT))


def _perm(n: int, pairs: Dict[int, int]) -> np.ndarray:
    m = np.zeros((n, n), dtype=CDT)
    for i in range(n):
        m[pairs.get(i, i), i] = 1
    return m


CNOT = controlled(X)  # gates.py:66-73
CZ = controlled(Z)  # gates.py:75-82
CY = controlled(Y)  # gates.py:84-91
SWAP = _perm(4, {1: 2, 2: 1})  # gates.py:93-100
TOFFOLI = controlled(CNOT)  # gates.py:103-114
FREDKIN = controlled(SWAP)  # gates.py:116-127

FIXED: Dict[str, np.ndarray] = {
    "i": I2,
    "x": X,
    "y": Y,
    "z": Z,
    "h": H,
    "t": T,
    "s": S,
    "td": T.conj().T,  # gates.py:974-976 (adjoint of t / s)
    "sd": S.conj().T,
    "wroot": WROOT,
    "cnot": CNOT,
    "cz": CZ,
    "swap": SWAP,
    "cy": CY,
    "ox": ocontrolled(X),  # gates.py:971-973
    "oy": ocontrolled(Y),
    "oz": ocontrolled(Z),
    "toffoli": TOFFOLI,
    "fredkin": FREDKIN,
}

ALIASES = {  # abstractcircuit.py:58-66
    "cx": "cnot",
    "cswap": "fredkin",
    "ccnot": "toffoli",
    "ccx": "toffoli",
    "unitary": "any",
    "sdg": "sd",
    "tdg": "td",
}


# ----------------------------------------------------------------------------------------
# parameterised matrices  (tensorcircuit/gates.py:463-865)
# ----------------------------------------------------------------------------------------
def _c(v: Any) -> complex:
    """num_to_tensor casts every parameter to the complex dtype (gates.py:225-232)."""
    return complex(v)


def m_phase(theta: Any = 0) -> np.ndarray:  # gates.py:463-482
    return np.diag([1.0, np.exp(1j * _c(theta))]).astype(CDT)


def m_u(theta: Any = 0, phi: Any = 0, lbd: Any = 0) -> np.ndarray:  # gates.py:509-542
    theta, phi, lbd = _c(theta), _c(phi), _c(lbd)
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array(
        [[c, -np.exp(1j * lbd) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lbd)) * c]],
        dtype=CDT,
    )


def m_r(theta: Any = 0, alpha: Any = 0, phi: Any = 0) -> np.ndarray:  # gates.py:545-573
    theta, alpha, phi = _c(theta), _c(alpha), _c(phi)
    return (
        np.cos(theta) * I2
        - 1j * np.cos(phi) * np.sin(alpha) * np.sin(theta) * X
        - 1j * np.sin(phi) * np.sin(alpha) * np.sin(theta) * Y
        - 1j * np.sin(theta) * np.cos(alpha) * Z
    )


def _rot(p: np.ndarray, theta: Any) -> np.ndarray:  # gates.py:579-636
    theta = _c(theta)
    return np.cos(theta / 2) * np.eye(p.shape[0], dtype=CDT) - 1j * np.sin(theta / 2) * p


def m_rx(theta: Any = 0) -> np.ndarray:
    return _rot(X, theta)


def m_ry(theta: Any = 0) -> np.ndarray:
    return _rot(Y, theta)


def m_rz(theta: Any = 0) -> np.ndarray:
    return _rot(Z, theta)


def m_iswap(theta: Any = 1.0) -> np.ndarray:  # gates.py:685-714
    theta = _c(theta)
    m = np.eye(4, dtype=CDT)
    c, s = np.cos(theta * np.pi / 2), np.sin(theta * np.pi / 2)
    m[1, 1] = m[2, 2] = c
    m[1, 2] = m[2, 1] = 1j * s
    return m


def m_cr(theta: Any = 0, alpha: Any = 0, phi: Any = 0) -> np.ndarray:  # gates.py:720-752
    return controlled(m_r(theta, alpha, phi))


def m_exp(unitary: Any, theta: Any) -> np.ndarray:  # gates.py:798-818: expm(-i theta U)
    u = _as_matrix(unitary)
    return scipy.linalg.expm(-1j * _c(theta) * u).astype(CDT)


def m_exp1(unitary: Any, theta: Any, half: bool = False) -> np.ndarray:  # gates.py:826-861
    u = _as_matrix(unitary)
    theta = _c(theta)
    if half:
        theta = theta / 2
    return np.cos(theta) * np.eye(u.shape[0], dtype=CDT) - 1j * np.sin(theta) * u


def _as_matrix(t: Any) -> np.ndarray:
    """reshapem: any [2]*2k tensor or square matrix -> 2^k x 2^k (abstract_backend.py:400-414)."""
    a = np.asarray(t, dtype=CDT)
    d = int(round(math.sqrt(a.size)))
    assert d * d == a.size, "gate tensor must have 4^k entries"
    return a.reshape(d, d)


def gate_matrix(name: str, **params: Any) -> np.ndarray:
    """Matrix U[out, in] (row-major over big-endian multi-indices) of a named gate.

    Name table: abstractcircuit.py:28-66; factories: gates.py:371-396, 949-982."""
    name = name.lower()
    name = ALIASES.get(name, name)
    if name in FIXED:
        return FIXED[name].copy()
    if name == "phase":
        return m_phase(**params)
    if name == "u":
        return m_u(**params)
    if name == "r":
        return m_r(**params)
    if name in ("rx", "ry", "rz"):
        return {"rx": m_rx, "ry": m_ry, "rz": m_rz}[name](**params)
    if name == "iswap":
        return m_iswap(**params)
    if name == "cr":
        return m_cr(**params)
    if name == "any":
        return _as_matrix(params["unitary"])
    if name == "exp":
        return m_exp(params.get("unitary", params.get("hermitian", params.get("hamiltonian"))), params["theta"])
    if name == "exp1":
        return m_exp1(
            params.get("unitary", params.get("hermitian", params.get("hamiltonian"))),
            params["theta"],
            params.get("half", False),
        )
    if name in ("rxx", "ryy", "rzz"):  # gates.py:863-865
        p = {"rxx": X, "ryy": Y, "rzz": Z}[name]
        return m_exp1(np.kron(p, p), params.get("theta", 0), half=True)
    if name in ("cu", "crx", "cry", "crz", "cphase"):  # gates.py:968-970
        return controlled(gate_matrix(name[1:], **params))
    if name in ("orx", "ory", "orz"):  # gates.py:971-973
        return ocontrolled(gate_matrix(name[1:], **params))
    raise ValueError("unknown gate %s" % name)


def multicontrol_matrix(unitary: Any, ctrl: Sequence[int]) -> np.ndarray:
    """Dense form of gates.multicontrol_gate (gates.py:868-942): U on the trailing legs iff
    the leading control legs read ``ctrl``; identity otherwise."""
    u = _as_matrix(unitary)
    if isinstance(ctrl, int):
        ctrl = [ctrl]
    nc = len(ctrl)
    d = u.shape[0]
    m = np.eye(d << nc, dtype=CDT)
    sel = 0
    for c in ctrl:
        sel = (sel << 1) | int(c)
    m[sel * d : (sel + 1) * d, sel * d : (sel + 1) * d] = u
    return m


# ----------------------------------------------------------------------------------------
# state evolution
# ----------------------------------------------------------------------------------------
def apply_gate(state: np.ndarray, u: np.ndarray, qubits: Sequence[int], n: int) -> np.ndarray:
    """psi'[.. o_j at q_j ..] = sum_i U[o, i] psi[.. i_j at q_j ..]

    (basecircuit.py:213-215: gate leg i+noe is wired to the front edge of qubit index[i],
    leg i becomes the new front; qubit 0 is the most significant index bit.)"""
    k = len(qubits)
    assert len(set(qubits)) == k  # basecircuit.py:143
    qubits = [q if q >= 0 else n + q for q in qubits]  # basecircuit.py:144
    t = state.reshape([2] * n)
    g = np.asarray(u, dtype=CDT).reshape([2] * (2 * k))
    t = np.tensordot(g, t, axes=(list(range(k, 2 * k)), list(qubits)))
    t = np.moveaxis(t, list(range(k)), list(qubits))
    return np.ascontiguousarray(t).reshape(-1)


class OracleCircuit:
    """Eager restatement of the tc.Circuit query surface used by the hot path."""

    def __init__(self, n: int, inputs: Optional[np.ndarray] = None):
        self.n = n
        if inputs is None:  # basecircuit.py:46-60
            self.psi = np.zeros(2**n, dtype=CDT)
            self.psi[0] = 1.0
            self.ntot = n
        else:  # circuit.py:86-96: 2^n or 2^(2n) entries (extra trailing legs)
            v = np.asarray(inputs, dtype=CDT).reshape(-1)
            m = int(round(math.log2(v.size)))
            assert m == n or m == 2 * n
            self.psi = v.copy()
            self.ntot = m
        self.ops: List[Tuple[str, Tuple[int, ...], Dict[str, Any]]] = []

    def gate(self, name: str, *qubits: int, **params: Any) -> "OracleCircuit":
        lname = ALIASES.get(name.lower(), name.lower())
        if lname == "multicontrol":
            u = multicontrol_matrix(params["unitary"], params.get("ctrl", 1))
        else:
            u = gate_matrix(lname, **params)
        assert u.shape[0] == 2 ** len(qubits), "gate %s arity mismatch" % name
        qubits = tuple(q if q >= 0 else self.n + q for q in qubits)
        self.psi = apply_gate(self.psi, u, qubits, self.ntot)
        self.ops.append((lname, qubits, params))
        return self

    def __getattr__(self, name: str):  # c.rx(0, theta=..), c.CNOT(0, 1), list broadcast
        lname = name.lower()
        if lname.startswith("_"):
            raise AttributeError(name)

        def f(*index: Any, **params: Any):
            # abstractcircuit.py:149-165: list-valued indices are zipped; list-valued params
            # are indexed per position when possible
            if isinstance(index[0], (int, np.integer)):
                return self.gate(lname, *index, **params)
            for i, ind in enumerate(zip(*index)):
                p = {}
                for k, v in params.items():
                    try:
                        p[k] = v[i]
                    except Exception:
                        p[k] = v
                self.gate(lname, *ind, **p)
            return self

        return f

    # -- queries ---------------------------------------------------------------------
    def state(self) -> np.ndarray:
        return self.psi.copy()

    wavefunction = state

    def expectation(self, *ops: Tuple[Any, Sequence[int]]) -> complex:
        """<psi| prod O_j |psi> for operators on disjoint sites (basecircuit.py:267-319,
        circuit.py:914-990); not normalised."""
        seen = set()
        phi = self.psi
        for op, idx in ops:
            if isinstance(idx, (int, np.integer)):
                idx = [idx]
            idx = [i if i >= 0 else self.n + i for i in idx]
            for i in idx:
                if i in seen:
                    raise ValueError("Cannot measure two operators in one index")
                seen.add(i)
            phi = apply_gate(phi, _as_matrix(op), idx, self.ntot)
        return complex(np.vdot(self.psi, phi))

    def expectation_ps(
        self,
        x: Optional[Sequence[int]] = None,
        y: Optional[Sequence[int]] = None,
        z: Optional[Sequence[int]] = None,
        ps: Optional[Sequence[int]] = None,
    ) -> complex:
        """abstractcircuit.py:1208-1288 (``ps`` overrides x/y/z) evaluated with the closed form
        of quantum.py:1461-1482."""
        x, y, z = resolve_ps(self.n, x, y, z, ps)
        return pauli_expectation(self.psi, self.n, x, y, z)

    def probability(self) -> np.ndarray:  # basecircuit.py:510-523
        return np.abs(self.psi) ** 2

    def sample_int(self, status: Sequence[float]) -> np.ndarray:
        return probability_sample(self.probability(), status)

    # -- Monte-Carlo trajectories (circuit.py:302-352, 473-744; basecircuit.py:824-857) -------
    def mid_measurement(self, index: int, keep: int = 0) -> int:  # circuit.py:302-347, no renorm
        proj = np.zeros((2, 2), dtype=CDT)
        proj[keep, keep] = 1.0
        self.psi = apply_gate(self.psi, proj, [index], self.ntot)
        return keep

    def unitary_kraus(self, kraus: Sequence[Any], *index: int, prob=None, status: float = 0.0) -> int:
        """circuit.py:473-565: the branch is floor-counted by the sign sum of 538-548."""
        ks = [_as_matrix(k) for k in kraus]
        if prob is None:
            prob = [float(np.real(np.trace(k.conj().T @ k)) / k.shape[0]) for k in ks]
            with np.errstate(divide="ignore", invalid="ignore"):
                ks = [k / np.sqrt(w + 0j) for k, w in zip(ks, prob)]
        cum = np.cumsum(np.real(np.asarray(prob, dtype=np.float64)))
        l = len(ks)
        r = int(sum(np.sign(status - cum[i]) for i in range(l - 1)) / 2.0 + (l - 1) / 2.0)
        self.psi = apply_gate(self.psi, ks[r], list(index), self.ntot)
        return r

    def general_kraus(self, kraus: Sequence[Any], *index: int, status: float = 0.0, with_prob: bool = False):
        """circuit.py:635-723: p_i = <K_i^dag K_i>, branch K_i / (sqrt(p_i) + 1e-10)."""
        ks = [_as_matrix(k) for k in kraus]
        prob = [float(np.real(self.expectation((k.conj().T @ k, list(index))))) for k in ks]
        new = [k / (np.sqrt(w) + 1e-10) for k, w in zip(ks, prob)]
        r = self.unitary_kraus(new, *index, prob=prob, status=status)
        return (r, prob) if with_prob else r

    def cond_measure(self, index: int, status: float = 0.0) -> int:  # basecircuit.py:824-857
        return self.general_kraus([np.diag([1.0, 0.0]), np.diag([0.0, 1.0])], index, status=status)

    # c.depolarizing / amplitudedamping / phasedamping / reset (circuit.py:725-744)
    def depolarizing(self, index: int, *, px: float, py: float, pz: float, status: float = 0.0) -> None:
        self.unitary_kraus(ch_depolarizing(px, py, pz), index, status=status)

    def amplitudedamping(self, index: int, *, gamma: float, p: float, status: float = 0.0) -> None:
        self.general_kraus(ch_amplitudedamping(gamma, p), index, status=status)

    def phasedamping(self, index: int, *, gamma: float, status: float = 0.0) -> None:
        self.general_kraus(ch_phasedamping(gamma), index, status=status)

    def reset(self, index: int, status: float = 0.0) -> None:
        self.general_kraus(ch_reset(), index, status=status)


def noisy_trajectory(c: Any, n: int, status: Sequence[float], theta: Sequence[float]) -> List[int]:
    """One Monte-Carlo trajectory recipe written against the shared Circuit surface (runs on the
    reference, on the oracle and on the product): two noisy entangling layers, a cond_measure,
    a reset, an explicit-prob 2-qubit unitary_kraus and a post-selection.  Returns the picks."""
    k = 0
    picks: List[int] = []
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
    picks.append(int(c.cond_measure(1, status=status[k])))
    c.reset(2, status=status[k + 1])
    zz = np.kron(Z, Z)
    xx = np.kron(X, X)
    picks.append(int(c.unitary_kraus([np.eye(4, dtype=CDT), zz, xx], 0, n - 1, prob=[0.5, 0.3, 0.2], status=status[k + 2])))
    c.mid_measurement(0, keep=1)
    return picks


NOISY_STATUS_LEN = lambda n: 6 * n + 3  # noqa: E731


# Kraus lists of channels.py:56-325 (plain matrices)
def ch_depolarizing(px: float, py: float, pz: float) -> List[np.ndarray]:
    return [np.sqrt(1 - px - py - pz + 0j) * I2, np.sqrt(px + 0j) * X, np.sqrt(py + 0j) * Y, np.sqrt(pz + 0j) * Z]


def ch_amplitudedamping(gamma: float, p: float) -> List[np.ndarray]:
    g, s = np.sqrt(gamma + 0j), np.sqrt(1 - gamma + 0j)
    return [
        np.sqrt(p + 0j) * np.array([[1, 0], [0, s]]),
        np.sqrt(p + 0j) * np.array([[0, g], [0, 0]]),
        np.sqrt(1 - p + 0j) * np.array([[s, 0], [0, 1]]),
        np.sqrt(1 - p + 0j) * np.array([[0, 0], [g, 0]]),
    ]


def ch_phasedamping(gamma: float) -> List[np.ndarray]:
    return [np.array([[1, 0], [0, np.sqrt(1 - gamma + 0j)]]), np.array([[0, 0], [0, np.sqrt(gamma + 0j)]])]


def ch_reset() -> List[np.ndarray]:
    return [np.array([[1, 0], [0, 0]], dtype=CDT), np.array([[0, 1], [0, 0]], dtype=CDT)]


def resolve_ps(n, x=None, y=None, z=None, ps=None):
    if ps is not None:  # quantum.py:1025-1044
        x = [i for i, p in enumerate(ps) if p == 1]
        y = [i for i, p in enumerate(ps) if p == 2]
        z = [i for i, p in enumerate(ps) if p == 3]
    fix = lambda l: [i if i >= 0 else n + i for i in (l or [])]
    x, y, z = fix(x), fix(y), fix(z)
    allq = x + y + z
    if len(set(allq)) != len(allq):
        raise ValueError("Cannot measure two operators in one index")
    return x, y, z


def pauli_masks(n: int, x: Iterable[int], y: Iterable[int], z: Iterable[int]) -> Tuple[int, int, int]:
    """(flip_mask, sign_mask, n_y): bit of qubit j is 1 << (n-1-j) (quantum.py:1439)."""
    mx = sum(1 << (n - 1 - j) for j in x)
    my = sum(1 << (n - 1 - j) for j in y)
    mz = sum(1 << (n - 1 - j) for j in z)
    return mx | my, my | mz, len(list(y))


def pauli_expectation(psi: np.ndarray, n: int, x, y, z) -> complex:
    """sum_r conj(psi_r) (-1)^{popc(r & (my|mz))} (-i)^{ny} psi_{r ^ (mx|my)}

    quantum.py:1461-1482: element (r, r^flip) of the Pauli string is (1-2e)(-i)^ny."""
    flip, sign, ny = pauli_masks(n, x, y, z)
    r = np.arange(2**n, dtype=np.int64)
    par = np.zeros(2**n, dtype=np.int64)
    t = r & sign
    for i in range(n):
        par ^= (t >> i) & 1
    vals = (1 - 2 * par) * ((-1j) ** (ny % 4))
    return complex(np.sum(np.conj(psi) * vals * psi[r ^ flip]))


def pauli_string_matrix(ps: Sequence[int], weight: complex = 1.0) -> np.ndarray:
    """Dense matrix of a Pauli string via the closed form (pins quantum.py:1461-1482 against
    tests/test_miscs.py:26-55)."""
    n = len(ps)
    x, y, z = resolve_ps(n, ps=ps)
    flip, sign, ny = pauli_masks(n, x, y, z)
    d = 2**n
    m = np.zeros((d, d), dtype=CDT)
    for r in range(d):
        e = bin(r & sign).count("1") & 1
        m[r, r ^ flip] = (1 - 2 * e) * ((-1j) ** (ny % 4)) * weight
    return m


# ----------------------------------------------------------------------------------------
# sampler  (abstract_backend.py:1124-1157, basecircuit.py:587-616)
# ----------------------------------------------------------------------------------------
def probability_sample(p: np.ndarray, status: Sequence[float], dtype: Any = np.float64) -> np.ndarray:
    """p/=sum(p); cdf=cumsum(p); r=cdf[-1]*(1-u); index=searchsorted(cdf, r, 'left').

    ``dtype`` is the real dtype the reference would run in (float32 for complex64 states);
    the parity tests evaluate the rule in float64 and allow differences only at CDF ties."""
    p = np.asarray(p, dtype=dtype)
    p = p / np.sum(p)
    cdf = np.cumsum(p)
    r = cdf[-1] * (1 - np.asarray(status).astype(dtype))
    return np.searchsorted(cdf, r, side="left").astype(np.int64)


def sample_cdf(p: np.ndarray) -> np.ndarray:
    p = np.asarray(p, dtype=np.float64)
    return np.cumsum(p / np.sum(p))


def sample_int2bin(sample: np.ndarray, n: int) -> np.ndarray:  # quantum.py:2104-2119
    sample = np.asarray(sample)
    return (sample[..., None] >> np.arange(n)[::-1]) % 2


def sample_bin2int(sample: np.ndarray, n: int) -> np.ndarray:  # quantum.py:2122-2134
    return np.sum(np.asarray(sample) * np.array([2**j for j in reversed(range(n))]), axis=-1)


def sample2count(sample: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:  # quantum.py:2137-2157
    return np.unique(np.asarray(sample), return_counts=True)


def count_s2d(srepr: Tuple[np.ndarray, np.ndarray], n: int) -> np.ndarray:  # quantum.py:2048-2064
    out = np.zeros(2**n, dtype=srepr[1].dtype)
    out[srepr[0]] = srepr[1]
    return out


def count_d2s(drepr: np.ndarray, eps: float = 1e-7) -> Tuple[np.ndarray, np.ndarray]:  # quantum.py:2070-2098
    drepr = np.asarray(drepr)
    idx = np.nonzero(np.abs(drepr) > eps)[0]
    return idx, drepr[idx]


def sample2all(sample: np.ndarray, n: int, format: str = "count_vector") -> Any:  # quantum.py:2324-2368
    sample = np.asarray(sample)
    if sample.ndim == 1:
        s_int, s_bin = sample, sample_int2bin(sample, n)
    elif sample.ndim == 2:
        s_int, s_bin = sample_bin2int(sample, n), sample
    else:
        raise ValueError("unrecognized tensor shape for sample")
    if format == "sample_int":
        return s_int
    if format == "sample_bin":
        return s_bin
    ct = sample2count(s_int)
    if format == "count_tuple":
        return ct
    if format == "count_vector":
        return count_s2d(ct, n)
    if format in ("count_dict_bin", "count_dict_int"):
        d = {int(i): int(j) for i, j in zip(*ct)}
        if format == "count_dict_int":
            return d
        return {bin(k)[2:].zfill(n): v for k, v in d.items()}
    raise ValueError("unsupported format %s for finite shots measurement" % format)


def spin_by_basis(n: int, m: int) -> np.ndarray:  # quantum.py:2371-2395
    """+1/-1 eigenvalue of Z on qubit m for every basis state (qubit 0 = MSB)."""
    r = np.arange(2**n)
    return 1 - 2 * ((r >> (n - 1 - m)) & 1)


def correlation_from_samples(index: Sequence[int], results: np.ndarray, n: int) -> float:
    """quantum.py:2398-2425: mean over shots of prod_{i in index} (1-2 b_i)."""
    results = np.asarray(results)
    if results.ndim == 1:
        results = sample_int2bin(results, n)
    r = 1 - 2 * results
    return float(np.mean(np.prod(r[:, list(index)], axis=-1)))


def correlation_from_counts(index: Sequence[int], results: np.ndarray) -> float:
    """quantum.py:2428-2451: results is a count_vector (any normalisation)."""
    results = np.asarray(results, dtype=np.float64)
    n = int(round(math.log2(results.size)))
    results = results / np.sum(results)
    for i in index:
        results = results * spin_by_basis(n, i)
    return float(np.sum(results))


# ----------------------------------------------------------------------------------------
# benchmark circuits of SURVEY.md section 8(d): (name, qubits, params) gate lists
# ----------------------------------------------------------------------------------------
GateList = List[Tuple[str, Tuple[int, ...], Dict[str, Any]]]


def hea_circuit(n: int, params: np.ndarray) -> GateList:
    """params [depth, 2, n]: per layer rx on all, rzz ladder, cnot ladder (config 1/3)."""
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
    """H on all, then per layer (rzz ladder; rx all): templates/blocks.py:141-152 shape.
    params [2*layers, n] (config 2)."""
    ops: GateList = [("h", (i,), {}) for i in range(n)]
    for l in range(params.shape[0] // 2):
        for i in range(n - 1):
            ops.append(("rzz", (i, i + 1), {"theta": float(params[2 * l, i])}))
        for i in range(n):
            ops.append(("rx", (i,), {"theta": float(params[2 * l + 1, i])}))
    return ops


def random_circuit(n: int, depth: int, seed: int) -> GateList:
    """Each layer: r(theta, alpha, phi) on every qubit then cnot on a random perfect matching
    (config 4/5; cf. gates.py:670-679 and examples/sample_benchmark.py:14-22)."""
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
    """X_i (weight -1) and Z_i Z_{i+1} (weight +1) as (weight, ps) -- examples/vqe_parallel_pmap.py:28-34."""
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


def run_gatelist(n: int, ops: GateList, inputs: Optional[np.ndarray] = None) -> OracleCircuit:
    c = OracleCircuit(n, inputs)
    for name, q, p in ops:
        c.gate(name, *q, **p)
    return c
