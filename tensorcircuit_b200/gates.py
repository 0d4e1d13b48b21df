"""Gate definitions of the hot path (host side, float64/complex128 numpy).

Mirrors the public surface of tensorcircuit/gates.py that ``Circuit`` relies on: the fixed
matrices (gates.py:31-127), ``Gate``, ``GateF`` / ``GateVF`` with ``adjoint / controlled /
ocontrolled / ided`` (gates.py:248-368), the parameterised gates (gates.py:463-865) and the
module-level registrations done by ``meta_gate`` / ``meta_vgate`` (gates.py:371-396, 949-982).

A gate tensor has shape ``[2]*2k`` with axes ``[out_0..out_{k-1}, in_0..in_{k-1}]``.  Matrices
are always built in complex128 on the host and cast to the state dtype when they are uploaded;
parameters may be :class:`~tensorcircuit_b200.batching.BatchArray` (inside ``backend.vmap``),
in which case the tensor carries a hidden leading batch axis."""

from __future__ import annotations

import sys
from functools import partial, reduce
from operator import mul
from typing import Any, Callable, List, Optional, Sequence, Union

import math

import numpy as np
import scipy.linalg

from .batching import BatchArray, is_batched

thismodule = sys.modules[__name__]

Tensor = Any
CDT = np.complex128

zero_state = np.array([1.0, 0.0], dtype=CDT)
one_state = np.array([0.0, 1.0], dtype=CDT)
plus_state = (zero_state + one_state) / np.sqrt(2)
minus_state = (zero_state - one_state) / np.sqrt(2)

_h_matrix = np.array([[1.0, 1.0], [1.0, -1.0]]) / np.sqrt(2)
_i_matrix = np.eye(2)
_x_matrix = np.array([[0.0, 1.0], [1.0, 0.0]])
_y_matrix = np.array([[0.0, -1j], [1j, 0.0]])
_z_matrix = np.diag([1.0, -1.0])
_s_matrix = np.diag([1.0, 1j])
_t_matrix = np.diag([1.0, np.exp(0.25j * np.pi)])
_wroot_matrix = np.array([[1, -(1 + 1j) / np.sqrt(2)], [(1 - 1j) / np.sqrt(2), 1]]) / np.sqrt(2)

_PAULI = {"i": _i_matrix, "x": _x_matrix, "y": _y_matrix, "z": _z_matrix}
for _a in "ixyz":
    for _b in "ixyz":
        setattr(thismodule, "_%s%s_matrix" % (_a, _b), np.kron(_PAULI[_a], _PAULI[_b]))


def _bd(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return scipy.linalg.block_diag(a, b)


_cnot_matrix = _bd(_i_matrix, _x_matrix)
_cz_matrix = _bd(_i_matrix, _z_matrix)
_cy_matrix = _bd(_i_matrix, _y_matrix)
_swap_matrix = np.eye(4)[[0, 2, 1, 3]]
_toffoli_matrix = _bd(np.eye(4), _cnot_matrix)
_fredkin_matrix = _bd(np.eye(4), _swap_matrix)


_NLEGS = {1 << k: k for k in range(0, 41)}


def _nlegs(size: int) -> int:
    n = _NLEGS.get(int(size))  # (a table: this sits on the path of every recorded gate)
    if n is None:
        raise ValueError("gate tensor size %d is not a power of two" % size)
    return n


def reshape2(t: Any) -> Any:
    """any tensor -> [2]*m  (abstract_backend.py:390-398)"""
    if is_batched(t):
        return t.reshape([2] * _nlegs(t.size))
    t = np.asarray(t)
    return t.reshape([2] * _nlegs(t.size))


def reshapem(t: Any) -> Any:
    """any tensor -> square matrix (abstract_backend.py:400-414)"""
    if is_batched(t):
        d = 2 ** (_nlegs(t.size) // 2)
        return t.reshape([d, d])
    t = np.asarray(t)
    d = 2 ** (_nlegs(t.size) // 2)
    return t.reshape(d, d)


class Gate:
    """Counterpart of tensorcircuit.gates.Gate (a tn.Node there, gates.py:138-177): just the
    tensor and a name -- there is no network to wire."""

    def __init__(self, tensor: Any, name: Optional[str] = None):
        if isinstance(tensor, Gate):
            tensor = tensor.tensor
        if not is_batched(tensor):
            tensor = np.asarray(tensor)
            if tensor.dtype.kind not in "c":
                tensor = tensor.astype(CDT)
            elif tensor.dtype != CDT:
                tensor = tensor.astype(CDT)
        elif tensor.dtype != CDT:
            tensor = tensor.astype(CDT)
        self.tensor = tensor
        self.name = name if name else "__unnamed_node__"
        self.kind: Optional[str] = None  # 'diag' / 'dense' when the constructor knows the zero pattern (fusion.matrix_kind otherwise)

    @property
    def batched(self) -> bool:
        return is_batched(self.tensor)

    def matrix(self) -> Any:
        return reshapem(self.tensor)

    def copy(self, conjugate: bool = False) -> "Gate":
        t = self.tensor.conj() if conjugate else self.tensor
        if not is_batched(t):
            t = np.array(t)
        return Gate(t, self.name)

    def __rmul__(self, lvalue: Any) -> "Gate":  # gates.py:130-135
        return Gate(lvalue * self.tensor)

    def __repr__(self) -> str:
        return "Gate(name=%r, tensor=%r)" % (self.name, self.tensor)


def num_to_tensor(*num: Any, dtype: Optional[str] = None) -> Any:
    """gates.py:180-238 -- everything becomes complex (host complex128)."""
    l = []
    for n in num:
        if is_batched(n):
            l.append(n.astype(CDT))
        else:
            l.append(np.asarray(n).astype(CDT))
    return l[0] if len(l) == 1 else l


array_to_tensor = num_to_tensor


class GateF:
    def __init__(self, m: Any, n: Optional[str] = None, ctrl: Optional[List[int]] = None):
        self.m = m
        self.n = n if n else "unknowngate"
        self.ctrl = ctrl

    def __call__(self, *args: Any, **kws: Any) -> Gate:
        return Gate(np.array(self.m, dtype=CDT), name=self.n)

    def adjoint(self) -> "GateF":
        m = self.__call__().tensor
        ma = reshapem(m).conj().T.reshape(m.shape)
        return GateF(ma, self.n + "d", self.ctrl)

    def ided(self, before: bool = True) -> "GateVF":
        def f(*args: Any, **kws: Any) -> Gate:
            u = reshapem(self.__call__(*args, **kws).tensor)
            iu = np.kron(np.eye(2), u) if before else np.kron(u, np.eye(2))
            return Gate(reshape2(iu), name=("ip" if before else "ia") + self.n)

        return GateVF(f, ("ip" if before else "ia") + self.n)

    def controlled(self) -> "GateVF":
        def f(*args: Any, **kws: Any) -> Gate:
            return Gate(reshape2(_blockdiag(None, reshapem(self.__call__(*args, **kws).tensor))), name="c" + self.n)

        return GateVF(f, "c" + self.n, [1] + (self.ctrl or []))

    def ocontrolled(self) -> "GateVF":
        def f(*args: Any, **kws: Any) -> Gate:
            return Gate(reshape2(_blockdiag(reshapem(self.__call__(*args, **kws).tensor), None)), name="o" + self.n)

        return GateVF(f, "o" + self.n, [0] + (self.ctrl or []))

    def __str__(self) -> str:
        return self.n

    __repr__ = __str__


def _blockdiag(a: Any, b: Any) -> Any:
    """[[a, 0], [0, b]] with None = identity of the other's size; batch aware."""
    ref = a if a is not None else b
    if is_batched(ref):
        s = ref.shape[-1]
        out = np.zeros((ref.batch, 2 * s, 2 * s), dtype=CDT)
        eye = np.eye(s)
        out[:, :s, :s] = eye if a is None else a.a
        out[:, s:, s:] = eye if b is None else b.a
        return BatchArray(out)
    s = ref.shape[-1]
    out = np.zeros((2 * s, 2 * s), dtype=CDT)
    out[:s, :s] = np.eye(s) if a is None else a
    out[s:, s:] = np.eye(s) if b is None else b
    return out


class GateVF(GateF):
    def __init__(self, f: Callable[..., Gate], n: Optional[str] = None, ctrl: Optional[List[int]] = None):
        self.f = f
        self.n = n if n else "unknowngate"
        self.ctrl = ctrl

    def __call__(self, *args: Any, **kws: Any) -> Gate:
        return self.f(*args, **kws)

    def adjoint(self) -> "GateVF":
        def f(*args: Any, **kws: Any) -> Gate:
            m = self.__call__(*args, **kws).tensor
            mm = reshapem(m)
            if is_batched(mm):
                ma = BatchArray(np.conj(np.swapaxes(mm.a, -1, -2))).reshape(m.shape)
            else:
                ma = mm.conj().T.reshape(m.shape)
            return Gate(ma, self.n + "d")

        return GateVF(f, self.n + "d", self.ctrl)


def meta_gate() -> None:
    """gates.py:371-396: ``x``, ``xgate``, ``x_gate`` ... from every ``_<name>_matrix``."""
    for name in dir(thismodule):
        if name.endswith("_matrix") and name.startswith("_"):
            n = name[1:-7]
            m = np.asarray(getattr(thismodule, name), dtype=CDT)
            temp = GateF(reshape2(m), n)
            for alias in (n + "gate", n + "_gate", n):
                setattr(thismodule, alias, temp)


meta_gate()
pauli_gates = [thismodule.i(), thismodule.x(), thismodule.y(), thismodule.z()]  # type: ignore


def matrix_for_gate(gate: Gate, tol: float = 1e-6) -> np.ndarray:
    t = np.array(reshapem(gate.tensor))
    t.real[abs(t.real) < tol] = 0.0
    t.imag[abs(t.imag) < tol] = 0.0
    return t


# ---------------------------------------------------------------------------------------------
# parameterised gates (gates.py:463-865)
# ---------------------------------------------------------------------------------------------
_E00 = np.array([[1, 0], [0, 0]], dtype=CDT)
_E01 = np.array([[0, 1], [0, 0]], dtype=CDT)
_E10 = np.array([[0, 0], [1, 0]], dtype=CDT)
_E11 = np.array([[0, 0], [0, 1]], dtype=CDT)


def _s(v: Any) -> Any:
    """scalar parameter -> complex scalar or batched scalar that broadcasts against matrices"""
    v = num_to_tensor(v)
    if is_batched(v):
        if v.size != 1:
            raise ValueError("gate parameter must be a scalar per batch element, got shape %r" % (tuple(v.shape),))
        return v.reshape([1, 1])
    if v.size != 1:
        # a vector angle would broadcast silently into a non-unitary matrix (n == 2) or fail later
        raise ValueError("gate parameter must be a scalar, got shape %r" % (tuple(np.shape(v)),))
    return v.reshape(())


def phase_gate(theta: float = 0) -> Gate:
    theta = _s(theta)
    return Gate(_E00 + np.exp(1.0j * theta) * _E11)


def u_gate(theta: float = 0, phi: float = 0, lbd: float = 0) -> Gate:
    theta, phi, lbd = _s(theta), _s(phi), _s(lbd)
    return Gate(
        np.cos(theta / 2) * _E00
        - np.exp(1.0j * lbd) * np.sin(theta / 2) * _E01
        + np.exp(1.0j * phi) * np.sin(theta / 2) * _E10
        + np.exp(1.0j * (phi + lbd)) * np.cos(theta / 2) * _E11
    )


def _rmat(theta: Any, alpha: Any, phi: Any) -> Any:
    if type(theta) in _REAL_SCALARS and type(alpha) in _REAL_SCALARS and type(phi) in _REAL_SCALARS:
        # plain real angles (the recording path of the random-circuit benchmark): scalar math, same formula
        ct, st = math.cos(float(theta)), math.sin(float(theta))
        sa, ca = math.sin(float(alpha)), math.cos(float(alpha))
        return (
            ct * _i_matrix
            - (1.0j * math.cos(float(phi)) * sa * st) * _x_matrix
            - (1.0j * math.sin(float(phi)) * sa * st) * _y_matrix
            - (1.0j * st * ca) * _z_matrix
        )
    theta, alpha, phi = _s(theta), _s(alpha), _s(phi)
    return (
        np.cos(theta) * _i_matrix
        - 1.0j * np.cos(phi) * np.sin(alpha) * np.sin(theta) * _x_matrix
        - 1.0j * np.sin(phi) * np.sin(alpha) * np.sin(theta) * _y_matrix
        - 1.0j * np.sin(theta) * np.cos(alpha) * _z_matrix
    )


def r_gate(theta: float = 0, alpha: float = 0, phi: float = 0) -> Gate:
    return Gate(_rmat(theta, alpha, phi))


_CS_BASIS: dict = {}
_CEYES: dict = {}


def _ceye(d: int) -> np.ndarray:
    e = _CEYES.get(d)
    if e is None:
        e = _CEYES[d] = np.eye(d, dtype=CDT)
    return e


def _cos_sin_batched(theta: Any, a: np.ndarray, b: np.ndarray, scale: float = 1.0) -> Any:
    """cos(scale theta) a - i sin(scale theta) b for a vmap batch of angles (the BatchArray as the user passed it,
    real or complex), in a handful of numpy calls on the raw [B] vector (the generic BatchArray arithmetic costs
    ~10 calls per gate)"""
    th = theta.a
    if th.size != th.shape[0]:
        raise ValueError("gate parameter must be a scalar per batch element, got shape %r" % (tuple(theta.shape),))
    th = th.reshape(-1)
    if th.dtype.kind == "c":
        if not th.imag.any():
            th = th.real
    elif th.dtype != np.float64:
        th = th.astype(np.float64)
    if scale != 1.0:
        th = th * scale
    c, s_ = np.cos(th), -1.0j * np.sin(th)
    # [B, 2] x [2, d*d]: one small matrix product builds the whole block (a zero entry of a and b stays an exact
    # zero; where both are non-zero the entries are +-1 / +-i, so the result is one exactly rounded addition).  One
    # strided update per non-zero entry was 8 passes over the array for an rzz: 240 us per gate at B = 1024.
    key = (id(a), id(b))
    ent = _CS_BASIS.get(key)
    if ent is None or ent[0] is not a or ent[1] is not b:
        ent = _CS_BASIS[key] = (a, b, np.ascontiguousarray(np.stack([np.asarray(a, dtype=CDT).reshape(-1), np.asarray(b, dtype=CDT).reshape(-1)])))
    coef = np.empty((th.shape[0], 2), dtype=CDT)
    coef[:, 0] = c
    coef[:, 1] = s_
    out = (coef @ ent[2]).reshape((th.shape[0],) + a.shape)
    return BatchArray(out)


def _rot(p: np.ndarray, theta: Any) -> Gate:
    if type(theta) in _REAL_SCALARS:  # plain real angle: scalar math instead of 0-d array arithmetic
        h = 0.5 * float(theta)
        return Gate(math.cos(h) * _i_matrix - (1.0j * math.sin(h)) * p)
    if is_batched(theta):
        g = Gate(_cos_sin_batched(theta, _i_matrix, p, 0.5))
        # (not the planner's half-cost class: with per-element matrices a merged block saves more -- one
        # matrix fetch and one dispatch per amplitude group -- than the halved FMAs of an unmerged rx / ry;
        # config 3: 131 ms merged, 151 ms unmerged)
        g.kind = "diag" if not (p[0, 1] or p[1, 0]) else "dense"
        return g
    theta = _s(theta)
    return Gate(np.cos(theta / 2.0) * _i_matrix - 1.0j * np.sin(theta / 2.0) * p)


def rx_gate(theta: float = 0) -> Gate:
    return _rot(_x_matrix, theta)


def ry_gate(theta: float = 0) -> Gate:
    return _rot(_y_matrix, theta)


def rz_gate(theta: float = 0) -> Gate:
    return _rot(_z_matrix, theta)


def rgate_theoretical(theta: float = 0, alpha: float = 0, phi: float = 0) -> Gate:
    theta, alpha, phi = complex(theta), complex(alpha), complex(phi)
    gen = np.sin(alpha) * np.cos(phi) * _x_matrix + np.sin(alpha) * np.sin(phi) * _y_matrix + np.cos(alpha) * _z_matrix
    return Gate(scipy.linalg.expm(-1.0j * theta * gen))


def random_single_qubit_gate() -> Gate:
    theta, alpha, phi = np.random.rand(3) * 2 * np.pi
    return r_gate(theta, alpha, phi)


def iswap_gate(theta: float = 1.0) -> Gate:
    theta = _s(theta)
    d1 = np.diag([1.0, 0, 0, 1.0]).astype(CDT)
    d2 = np.diag([0, 1.0, 1.0, 0]).astype(CDT)
    od = np.zeros((4, 4), dtype=CDT)
    od[1, 2] = od[2, 1] = 1.0
    return Gate(reshape2(d1 + np.cos(theta * np.pi / 2) * d2 + 1.0j * np.sin(theta * np.pi / 2) * od))


def cr_gate(theta: float = 0, alpha: float = 0, phi: float = 0) -> Gate:
    return Gate(reshape2(_blockdiag(None, _rmat(theta, alpha, phi))))


def random_two_qubit_gate() -> Gate:
    from scipy.stats import unitary_group

    return Gate(reshape2(unitary_group.rvs(dim=4)), name="R2Q")


def any_gate(unitary: Tensor, name: str = "any") -> Gate:
    if isinstance(unitary, Gate):
        return unitary
    if hasattr(unitary, "__array__") and not is_batched(unitary):
        unitary = np.asarray(unitary)
    return Gate(reshape2(unitary), name=name)


def _alias_unitary(kws: dict) -> dict:
    for a in ("hermitian", "hamiltonian"):  # arg_alias, gates.py:783, 821
        if a in kws:
            kws["unitary"] = kws.pop(a)
    return kws


def exponential_gate(unitary: Tensor = None, theta: float = None, name: str = "none", **kws: Any) -> Gate:
    kws = _alias_unitary(kws)
    unitary = kws.get("unitary", unitary)
    if is_batched(theta) or is_batched(unitary):
        raise NotImplementedError("exp gate inside vmap: use exp1 for involutory generators")
    u = reshapem(np.asarray(unitary, dtype=CDT))
    return Gate(reshape2(scipy.linalg.expm(-1.0j * complex(np.asarray(theta)) * u)), name="exp-" + name)


exp_gate = exponential_gate


_EYES: dict = {}
_EXP1_CACHE: dict = {}
_EXP1_BATCHED: dict = {}
_REAL_SCALARS = (float, int, np.float64, np.float32)


def exponential_gate_unity(unitary: Tensor = None, theta: float = None, half: bool = False, name: str = "none", **kws: Any) -> Gate:
    kws = _alias_unitary(kws)
    unitary = kws.get("unitary", unitary)
    if type(theta) in _REAL_SCALARS and isinstance(unitary, np.ndarray):
        # plain real angle (the common case on the recording path): scalar math, cached generator
        ent = _EXP1_CACHE.get(id(unitary))
        if ent is None or ent[0] is not unitary:
            um = np.asarray(unitary, dtype=CDT)
            d = 2 ** (_nlegs(um.size) // 2)
            ent = _EXP1_CACHE[id(unitary)] = (unitary, um.reshape(d, d), np.eye(d, dtype=CDT), [2] * _nlegs(um.size))
        t = float(theta) * (0.5 if half is True else 1.0)
        mat = math.cos(t) * ent[2] - (1.0j * math.sin(t)) * ent[1]
        return Gate(mat.reshape(ent[3]), name="exp1-" + name)
    if is_batched(theta):
        ent = _EXP1_BATCHED.get(id(unitary)) if isinstance(unitary, np.ndarray) else None
        if ent is None or ent[0] is not unitary:
            um = np.asarray(unitary, dtype=CDT)
            n = _nlegs(um.size)
            d = 2 ** (n // 2)
            um = um.reshape(d, d)
            ent = (unitary, um, _ceye(d), [2] * n, "diag" if not um[~np.eye(d, dtype=bool)].any() else "dense")
            if isinstance(unitary, np.ndarray):  # stable objects for the basis cache of _cos_sin_batched
                _EXP1_BATCHED[id(unitary)] = ent
        mat = _cos_sin_batched(theta, ent[2], ent[1], 0.5 if half is True else 1.0)
        g = Gate(mat.reshape(ent[3]), name="exp1-" + name)
        g.kind = ent[4]
        return g
    u = np.asarray(unitary, dtype=CDT)
    n = _nlegs(u.size)
    d = 2 ** (n // 2)
    theta = _s(theta)
    if half is True:
        theta = theta / 2.0
    eye = _EYES.get(d)
    if eye is None:
        eye = _EYES.setdefault(d, np.eye(d))
    mat = np.cos(theta) * eye - 1.0j * np.sin(theta) * u.reshape(d, d)
    return Gate(reshape2(mat), name="exp1-" + name)


exp1_gate = exponential_gate_unity

rzz_gate = partial(exp1_gate, unitary=thismodule._zz_matrix, half=True)  # type: ignore
rxx_gate = partial(exp1_gate, unitary=thismodule._xx_matrix, half=True)  # type: ignore
ryy_gate = partial(exp1_gate, unitary=thismodule._yy_matrix, half=True)  # type: ignore


def multicontrol_gate(unitary: Tensor, ctrl: Union[int, Sequence[int]] = 1) -> Gate:
    """Dense form of the reference's MPO multi-control gate (gates.py:868-942): ``unitary`` on
    the trailing legs iff the leading control legs read ``ctrl``."""
    if isinstance(unitary, Gate):
        unitary = unitary.tensor
    u = reshapem(np.asarray(unitary, dtype=CDT))
    if isinstance(ctrl, (int, np.integer)):
        ctrl = [int(ctrl)]
    d = u.shape[0]
    m = np.eye(d << len(ctrl), dtype=CDT)
    sel = reduce(lambda a, c: (a << 1) | int(round(float(np.real(c)))), ctrl, 0)
    m[sel * d : (sel + 1) * d, sel * d : (sel + 1) * d] = u
    return Gate(reshape2(m), name="multicontrol")


def mpo_gate(mpo: Any, name: str = "mpo") -> Any:
    raise NotImplementedError("MPO-form operators are outside the statevector hot path")


def meta_vgate() -> None:
    """gates.py:949-982"""
    for f in ["r", "u", "rx", "ry", "rz", "phase", "iswap", "any", "exp", "exp1", "cr", "rzz", "rxx", "ryy"]:
        for funcname in [f, f + "gate"]:
            setattr(thismodule, funcname, GateVF(getattr(thismodule, f + "_gate"), f))
    for f in ["cu", "crx", "cry", "crz", "cphase"]:
        for funcname in [f, f + "gate"]:
            setattr(thismodule, funcname, getattr(thismodule, f[1:]).controlled())
    for f in ["ox", "oy", "oz", "orx", "ory", "orz"]:
        for funcname in [f, f + "gate"]:
            setattr(thismodule, funcname, getattr(thismodule, f[1:]).ocontrolled())
    for f in ["sd", "td"]:
        for funcname in [f, f + "gate"]:
            setattr(thismodule, funcname, getattr(thismodule, f[:-1]).adjoint())
    for f in ["multicontrol", "mpo"]:
        for funcname in [f, f + "gate"]:
            setattr(thismodule, funcname, GateVF(getattr(thismodule, f + "_gate"), f))


meta_vgate()
