"""``value_and_grad`` / ``vvag`` for loss functions built from circuit expectation values.

The reference differentiates through its backend's autodiff (jax_backend.py:668-776).  The
engine has no autodiff, but an expectation value is a *quadratic form in every gate matrix*:
with psi(M_j) linear in the matrix of gate j,

    E(M_j + D) - E(M_j - D) = 4 Re <psi| H |d psi>,      d psi = psi with gate j replaced by D,

so for D = dM_j/d theta the derivative of E through gate j is exactly (E+ - E-)/2 -- a shift
rule that holds for ANY gate (no generator assumption, no truncation error), needs only forward
simulations with non-unitary matrices, and therefore runs on the existing kernels: all shifted
circuits of a query are evaluated as ONE vmap-style batch (batched matrices for the shifted
gates only).  Cost: two batch elements per (parameter, dependent gate) pair, like a
parameter-shift gradient.  That path is the FALLBACK.  The default is the adjoint-state sweep
(``_adjoint_contrib``, O(1) simulations): with lambda = H psi and both states swept backwards
through the circuit, the derivative through gate j is 2 Re <lambda_j| dM_j M_j^+ |psi_j> -- one
forward simulation, one backward sweep over a two-row state, one reduction launch per group of
commuting taps (csrc/sparse.cu).  It needs unitary gates and derivative-carrying gates on <= 2 qubits;
anything else goes through the shift rule.

What is differentiated is the function itself, not a restricted form of it:
  * dM_j/d theta_k comes from re-running the *recording* of ``f`` (no device work) at
    theta_k +- h and differencing the recorded gate matrices (central difference on smooth
    2x2 / 4x4 matrices in float64: error ~1e-9);
  * the host arithmetic that turns expectation values into the loss is differentiated the same
    way: ``f`` is replayed with one stored expectation value nudged at a time (exact when the
    loss is linear in the expectation values, as in every energy function).
Limits: real parameters; queries are expectation values (``expectation_ps``, ``expectation``,
the measurement templates); circuits start from |0..0>; the structure of ``f`` must not depend on
the parameter values (the same restriction ``jit`` has in the reference)."""

from __future__ import annotations

import hashlib
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import circuit as _circuit
from . import engine as _engine
from .batching import BatchArray

_H_PARAM = 1e-4   # central-difference step for d(gate matrix)/d(parameter)
_H_VALUE = 1e-3   # nudge of a stored expectation value when differentiating the host arithmetic
_MAX_BATCH_AMPS = 1 << 28


class _Query:
    def __init__(self, circ: Any, fl: Sequence[int], sg: Sequence[int], ny: Sequence[int], need_key: bool = True):
        self.nqubits = circ._nqubits
        self.circ = circ  # strong reference: the circuit (and its id) outlives the trace
        # matrix [D, D], or [B, D, D] when the recording itself was run on a batch of nudged parameters
        self.ops = [(tuple(op.qubits), np.array(op.matrix.a if isinstance(op.matrix, BatchArray) else np.asarray(op.matrix), dtype=np.complex128))
                    for op in circ._ops]
        # queries on the same gate sequence share their shifted simulations.  The key is the gate
        # CONTENT (qubits + matrix bytes), not the circuit object: sample_expectation_ps appends and
        # removes basis rotations on one circuit, and ids are reused after garbage collection.
        # (only the queries of the device run are grouped; replayed ones are matched by position and their
        # matrices carry a batch axis of all nudged parameters -- hashing those was 20 % of a gradient call)
        self.key = None
        if need_key:
            h = hashlib.blake2b(digest_size=16)
            for qubits, M in self.ops:
                h.update(repr((qubits, M.shape)).encode())
                h.update(np.ascontiguousarray(M).tobytes())
            self.key = (self.nqubits, len(self.ops), h.digest())
        self.fl, self.sg, self.ny = list(fl), list(sg), list(ny)
        self.coo: Any = None  # a device-resident sparse operator instead of Pauli strings (one value)
        self.values: Optional[np.ndarray] = None  # complex [nterms]

    @property
    def nterms(self) -> int:
        return 1 if self.coo is not None else len(self.fl)


class _Trace:
    """Context shared by the recording / replaying state stand-ins."""

    def __init__(self, replay: Optional[List[np.ndarray]] = None, value_batch: Optional[int] = None):
        self.queries: List[_Query] = []
        self.replay = replay            # per query: complex [nterms], or [value_batch, nterms]
        self.value_batch = value_batch  # replay a whole batch of nudged expectation values at once


class _RecordingState:
    """Wraps the real device state: every expectation query is logged with the circuit that asked."""

    def __init__(self, real: Any, trace: _Trace):
        self.__dict__["_real"] = real
        self.__dict__["_trace"] = trace
        self.__dict__["_circ"] = None

    def __getattr__(self, name: str) -> Any:
        return getattr(self._real, name)

    def __setattr__(self, name: str, value: Any) -> None:
        if name in ("_circ",):
            self.__dict__[name] = value
        else:
            setattr(self._real, name, value)

    def expectation_terms(self, fl: Sequence[int], sg: Sequence[int], ny: Sequence[int]) -> np.ndarray:
        r = self._real.expectation_terms(fl, sg, ny)
        if self._real.batch != 1:
            raise NotImplementedError("value_and_grad of a function that vmaps internally")
        q = _Query(self._circ, fl, sg, ny)
        q.values = np.array(r[0], dtype=np.complex128)
        self._trace.queries.append(q)
        return r


def _record_coo(self: Any, op: Any) -> np.ndarray:
    r = self._real.coo_expectation(op)
    if self._real.batch != 1:
        raise NotImplementedError("value_and_grad of a function that vmaps internally")
    q = _Query(self._circ, [], [], [])
    q.coo = op  # (same key as the Pauli queries on this circuit: they share one sweep)
    q.values = np.array([r[0]], dtype=np.complex128)
    self._trace.queries.append(q)
    return r


_RecordingState.coo_expectation = _record_coo  # type: ignore[attr-defined]


class _ReplayState:
    """No device at all: gates are only recorded by the Circuit, queries return stored values
    (tiled over the batch when the recording runs on a batch of nudged parameters)."""

    def __init__(self, nbits: int, dtype: str, batch: int, trace: _Trace):
        self.nbits, self.dtype, self._trace, self._circ, self.batch = nbits, dtype, trace, None, int(batch)

    def init_zero(self) -> None:
        pass

    def load(self, src: Any) -> None:
        raise NotImplementedError("value_and_grad: circuits with inputs= are not differentiable here yet")

    def apply_planned(self, blocks: Any) -> int:
        return 0

    def apply_blocks(self, blocks: Any) -> None:
        pass

    def expectation_terms(self, fl: Sequence[int], sg: Sequence[int], ny: Sequence[int]) -> np.ndarray:
        i = len(self._trace.queries)
        q = _Query(self._circ, fl, sg, ny, need_key=False)
        self._trace.queries.append(q)
        if self._trace.replay is None or i >= len(self._trace.replay) or np.shape(self._trace.replay[i])[-1] != len(fl):
            raise RuntimeError("value_and_grad: the structure of the function changed between evaluations")
        v = np.asarray(self._trace.replay[i], dtype=np.complex128)
        return v if v.ndim == 2 else np.tile(v[None, :], (self.batch, 1))

    def coo_expectation(self, op: Any) -> np.ndarray:
        i = len(self._trace.queries)
        q = _Query(self._circ, [], [], [], need_key=False)
        q.coo = op
        self._trace.queries.append(q)
        if self._trace.replay is None or i >= len(self._trace.replay) or np.shape(self._trace.replay[i])[-1] != 1:
            raise RuntimeError("value_and_grad: the structure of the function changed between evaluations")
        v = np.asarray(self._trace.replay[i], dtype=np.complex128)
        return v[:, 0] if v.ndim == 2 else np.tile(v[None, :], (self.batch, 1))[:, 0]

    def __getattr__(self, name: str) -> Any:
        raise NotImplementedError("value_and_grad: %s() is not differentiable (only expectation values are)" % name)


def _run(f: Callable[..., Any], args: Sequence[Any], kws: Dict[str, Any], trace: _Trace, replay: bool) -> Any:
    """Call ``f`` with the Circuit class bound to a recording (device) or replaying (host) state."""
    real_cls = _engine.DeviceState

    def factory(nbits: int, dtype: str = "complex64", batch: int = 1, **kw: Any) -> Any:
        if replay:
            return _ReplayState(nbits, dtype, batch, trace)
        return _RecordingState(real_cls(nbits, dtype, batch, **kw), trace)

    old_hook = _circuit._STATE_HOOK
    _engine.DeviceState = factory  # type: ignore[assignment]

    def hook(circ: Any, st: Any) -> None:
        st._circ = circ
        if replay and trace.value_batch:  # queries must hand BatchArrays to the host arithmetic
            circ._batch = trace.value_batch
            st.batch = trace.value_batch

    _circuit._STATE_HOOK = hook
    try:
        return f(*args, **kws)
    finally:
        _engine.DeviceState = real_cls  # type: ignore[assignment]
        _circuit._STATE_HOOK = old_hook


def _scalar(out: Any, has_aux: bool) -> Tuple[float, Any]:
    aux = None
    if has_aux:
        out, aux = out[0], out[1]
    v = np.asarray(out)
    if v.size != 1:
        raise ValueError("value_and_grad needs a scalar loss")
    return float(np.real(v.reshape(-1)[0])), aux


def _shift_batch(q: _Query, base: _Query, shifts: List[Tuple[int, np.ndarray]], dtype: str) -> np.ndarray:
    """E+ - E- for every (gate j, D) in ``shifts`` on query ``q``: complex [nshifts, nterms], each the
    difference of the query's expectation values with gate j replaced by M_j + D and M_j - D."""
    n = q.nqubits
    out = np.zeros((len(shifts), len(q.fl)), dtype=np.complex128)
    per = max(1, (_MAX_BATCH_AMPS >> n) // 2)  # shifts per batched run (two states each)
    for c0 in range(0, len(shifts), per):
        chunk = shifts[c0 : c0 + per]
        B = 2 * len(chunk)
        touched: Dict[int, np.ndarray] = {}
        for b, (j, D) in enumerate(chunk):
            if j not in touched:
                touched[j] = np.broadcast_to(base.ops[j][1], (B,) + base.ops[j][1].shape).copy()
            touched[j][2 * b] = base.ops[j][1] + D
            touched[j][2 * b + 1] = base.ops[j][1] - D
        c = _circuit.Circuit(n)
        for j, (qubits, M) in enumerate(base.ops):
            c.any(*qubits, unitary=BatchArray(touched[j]) if j in touched else M)
        if c._batch is None:  # nothing batched (cannot happen with a non-empty chunk)
            continue
        st = c._ensure_state()
        r = np.asarray(st.expectation_terms(q.fl, q.sg, q.ny), dtype=np.complex128)  # [B, nterms]
        out[c0 : c0 + len(chunk)] = r[0::2] - r[1::2]
    return out


_ADJOINT = True      # False: always use the shift rule (tests compare the two)
ADJOINT_STATS = {"sweeps": 0, "flushes": 0, "tap_launches": 0, "fallbacks": 0}


def _embed(m: np.ndarray, qubits: Sequence[int], union: Sequence[int]) -> np.ndarray:
    """operator ``m`` on ``qubits`` (caller's order, big-endian) as a matrix on the ascending qubit list ``union``"""
    from .fusion import embed_apply

    return embed_apply(np.eye(1 << len(union), dtype=np.complex128), list(union), np.asarray(m, dtype=np.complex128), list(qubits))


def _commute(a: np.ndarray, qa: Sequence[int], b: np.ndarray, qb: Sequence[int]) -> bool:
    if not set(qa) & set(qb):
        return True
    da = np.count_nonzero(a - np.diag(np.diagonal(a))) == 0
    db = np.count_nonzero(b - np.diag(np.diagonal(b))) == 0
    if da and db:
        return True
    u = sorted(set(qa) | set(qb))
    A, B = _embed(a, qa, u), _embed(b, qb, u)
    return bool(np.abs(A @ B - B @ A).max() < 1e-10)


def _sorted_local(m: np.ndarray, qubits: Sequence[int], n: int) -> Tuple[List[int], np.ndarray]:
    """(ascending amplitude-index bits, matrix with index bit i <-> bits[i]) of an operator given on
    ``qubits`` in the caller's order"""
    qs = sorted(qubits)
    return [n - 1 - q for q in reversed(qs)], _embed(m, qubits, qs)


def _adjoint_contrib(q: _Query, weights: np.ndarray, lst: List[Tuple[int, int, np.ndarray]], dtype: str) -> Optional[np.ndarray]:
    """d(sum_t weights_t Re E_t)/d theta through gate j with dM_j = D, for every (k, j, D) of ``lst``, by
    one backward sweep.  None when the circuit is outside the sweep's scope (caller falls back)."""
    n = q.nqubits
    if n > 31:
        return None
    Mdag: List[np.ndarray] = []
    for qubits, M in q.ops:
        if M.ndim != 2 or np.abs(M @ M.conj().T - np.eye(M.shape[0])).max() > 1e-9:
            return None  # non-unitary gate (Kraus branch, projector): cannot be undone by its adjoint
        Mdag.append(M.conj().T)
    taps: Dict[int, List[int]] = {}
    for idx, (_, j, D) in enumerate(lst):
        if len(q.ops[j][0]) > 2:
            return None
        taps.setdefault(j, []).append(idx)
    st2 = _engine.DeviceState(n, dtype, 2)
    circ = q.circ
    src = getattr(circ, "_state", None)
    if src is None or getattr(circ, "_applied", -1) != len(q.ops) or len(circ._ops) != len(q.ops) or getattr(src, "batch", 0) != 1:
        c = _circuit.Circuit(n)  # the recorded circuit has moved on: one more forward simulation
        for qubits, M in q.ops:
            c.any(*qubits, unitary=M)
        src = c._ensure_state()
    st2.copy_row_from(0, src, 0)
    coef = [w * (-1j) ** (int(y) & 3) for w, y in zip(weights, q.ny)]
    keep = [t for t in range(len(coef)) if coef[t] != 0]
    coos = [(op, w) for op, w in getattr(q, "coos", []) if w != 0]
    if not keep and not coos:
        return np.zeros(len(lst))
    for op, _ in coos:
        op.csr()
        if not getattr(op, "hermitian", True):
            return None  # 2 Re <H psi| d psi> is the derivative only for a Hermitian H
    if keep:
        st2.apply_pauli_sum_rows(0, 1, [q.fl[t] for t in keep], [q.sg[t] for t in keep], [coef[t] for t in keep])
    for ci, (op, w) in enumerate(coos):  # lambda += w H psi, row by row of the CSR form
        st2.apply_csr_rows(0, 1, op, coef=w, accumulate=bool(keep) or ci > 0)
    out = np.zeros(len(lst))
    pending_undo: List[Tuple[Tuple[int, ...], np.ndarray]] = []
    on_qubit: Dict[int, List[int]] = {}  # qubit -> indices into pending_undo (only those can fail to commute with a tap)
    pending_taps: List[Tuple[int, List[int], np.ndarray]] = []
    fuser = _circuit.Circuit(n)

    def flush() -> None:
        if pending_taps:
            vals = st2.transition_local(1, 0, [(bits, G) for _, bits, G in pending_taps])
            for (idx, _, _), v in zip(pending_taps, vals):
                out[idx] = 2.0 * float(np.real(v))
            ADJOINT_STATS["tap_launches"] += 1
            pending_taps.clear()
        if pending_undo:
            from .fusion import GateOp

            blocks = fuser._fuse([GateOp(tuple(qs), m) for qs, m in pending_undo], n)
            if hasattr(st2, "apply_planned") and fuser.use_passes:
                st2.apply_planned(blocks)
            else:
                st2.apply_blocks(blocks)
            ADJOINT_STATS["flushes"] += 1
            pending_undo.clear()
            on_qubit.clear()

    jmin = min(taps)
    for j in range(len(q.ops) - 1, jmin - 1, -1):
        qubits, M = q.ops[j]
        for idx in taps.get(j, []):
            G = lst[idx][2] @ Mdag[j]
            near = sorted({i for qb in qubits for i in on_qubit.get(qb, ())})
            if any(not _commute(G, qubits, pending_undo[i][1], pending_undo[i][0]) for i in near):
                flush()  # taps registered so far see the current states; then the states move on
            bits, Gs = _sorted_local(G, qubits, n)
            pending_taps.append((idx, bits, Gs))
        if j > jmin:
            for qb in qubits:
                on_qubit.setdefault(qb, []).append(len(pending_undo))
            pending_undo.append((tuple(qubits), Mdag[j]))
    pending_undo.clear()  # nobody needs the states below the lowest tap
    flush()
    ADJOINT_STATS["sweeps"] += 1
    return out


def value_and_grad(f: Callable[..., Any], argnums: Union[int, Sequence[int]] = 0, has_aux: bool = False) -> Callable[..., Any]:
    single = isinstance(argnums, int)
    nums: Tuple[int, ...] = (argnums,) if single else tuple(argnums)  # type: ignore[assignment]

    def wrapper(*args: Any, **kws: Any) -> Any:
        args = list(args)
        for i in nums:
            a = np.asarray(args[i])
            if np.iscomplexobj(a):
                if np.abs(a.imag).max() > 0:
                    raise NotImplementedError("value_and_grad with respect to complex parameters")
                a = a.real
            args[i] = np.array(a, dtype=np.float64)
        # 1. the value, on the device, with every expectation query logged
        base = _Trace()
        out = _run(f, args, kws, base, replay=False)
        value, aux = _scalar(out, has_aux)
        stored = [q.values for q in base.queries]

        def replay(a: Sequence[Any], values: List[np.ndarray]) -> Tuple[float, _Trace]:
            t = _Trace(replay=values)
            v, _ = _scalar(_run(f, a, kws, t, replay=True), has_aux)
            return v, t

        # 2. d loss / d (expectation value), by nudging one stored value at a time
        dl_de = [np.zeros(len(v)) for v in stored]
        nudged = False
        try:  # all 2m nudges in ONE host pass: the expectation values become BatchArrays
            m = sum(len(v) for v in stored)
            if m:
                vb = [np.tile(v[None, :], (2 * m, 1)) for v in stored]
                steps = []
                r = 0
                for qi, v in enumerate(stored):
                    for t in range(len(v)):
                        st_ = _H_VALUE * max(1.0, abs(v[t]))
                        vb[qi][2 * r, t] += st_
                        vb[qi][2 * r + 1, t] -= st_
                        steps.append(st_)
                        r += 1
                tr = _Trace(replay=vb, value_batch=2 * m)
                outb = _run(f, args, kws, tr, replay=True)
                if has_aux:
                    outb = outb[0]
                if len(tr.queries) != len(stored):
                    raise RuntimeError("structure")
                if isinstance(outb, BatchArray):
                    lv = np.real(np.asarray(outb.a, dtype=np.complex128)).reshape(2 * m)
                    d = (lv[0::2] - lv[1::2]) / (2 * np.asarray(steps))
                    r = 0
                    for qi, v in enumerate(stored):
                        dl_de[qi][:] = d[r : r + len(v)]
                        r += len(v)
                    nudged = True  # otherwise: fall through to one replay per nudged value
        except (TypeError, ValueError, NotImplementedError, RuntimeError, AttributeError, IndexError):
            dl_de = [np.zeros(len(v)) for v in stored]
        for qi, v in enumerate(stored if not nudged else []):
            for t in range(len(v)):
                step = _H_VALUE * max(1.0, abs(v[t]))
                up = [w.copy() for w in stored]
                dn = [w.copy() for w in stored]
                up[qi][t] += step
                dn[qi][t] -= step
                dl_de[qi][t] = (replay(args, up)[0] - replay(args, dn)[0]) / (2 * step)
        # 3. per parameter: direct dependence of the host arithmetic + the gates it moves
        dtype = _circuit.cons.dtypestr
        grads = []
        # queries asked of the same circuit prefix (a loop of expectation_ps calls) form one group
        groups: Dict[Any, List[int]] = {}
        for qi, qb in enumerate(base.queries):
            groups.setdefault(qb.key, []).append(qi)
        leader = {qi: members[0] for members in groups.values() for qi in members}
        # the derivative-carrying gates of ALL differentiated arguments go through one sweep per circuit:
        # per group leader (argument position, parameter k, gate j, D)
        all_pending: List[List[Tuple[int, int, int, np.ndarray]]] = [[] for _ in base.queries]
        gflat: List[np.ndarray] = []
        for pos, i in enumerate(nums):
            theta = args[i]
            g = np.zeros(theta.size)
            gflat.append(g)
            flat = theta.reshape(-1)
            pending: List[List[Tuple[int, int, np.ndarray]]] = [[] for _ in base.queries]  # per group leader: (k, gate j, D)
            swept = False
            try:  # all 2P nudged recordings in ONE host pass, the way vmap runs a batch of parameters
                P = theta.size
                hs = _H_PARAM * np.maximum(1.0, np.abs(flat))
                tb = np.repeat(flat[None, :], 2 * P, axis=0)
                tb[2 * np.arange(P), np.arange(P)] += hs
                tb[2 * np.arange(P) + 1, np.arange(P)] -= hs
                ab = list(args)
                ab[i] = BatchArray(tb.reshape((2 * P,) + theta.shape))
                tr = _Trace(replay=stored)
                outb = _run(f, ab, kws, tr, replay=True)
                if has_aux:
                    outb = outb[0]
                if len(tr.queries) != len(base.queries):
                    raise RuntimeError("structure")
                if isinstance(outb, BatchArray):
                    lv = np.real(np.asarray(outb.a, dtype=np.complex128)).reshape(2 * P)
                    g += (lv[0::2] - lv[1::2]) / (2 * hs)
                for qi, (qt, qb) in enumerate(zip(tr.queries, base.queries)):
                    if len(qt.ops) != len(qb.ops):
                        raise RuntimeError("structure")
                    if leader[qi] != qi or not any(np.any(dl_de[m]) for m in groups[qb.key]):
                        continue
                    # gates that see the parameters carry a [2P, d, d] matrix; all gates of one width in one pass
                    by_shape: Dict[Any, List[int]] = {}
                    for j, (_, Mb) in enumerate(qt.ops):
                        if Mb.ndim == 3:
                            by_shape.setdefault(Mb.shape, []).append(j)
                    found: List[Tuple[int, int, np.ndarray]] = []
                    for js in by_shape.values():
                        Mall = np.stack([qt.ops[j][1] for j in js])  # [gates, 2P, d, d]
                        Dall = (Mall[:, 0::2] - Mall[:, 1::2]) / (2 * hs)[None, :, None, None]
                        hit = np.abs(Dall).reshape(len(js), P, -1).max(axis=2) > 1e-12
                        for gi, k in zip(*np.nonzero(hit)):
                            found.append((js[gi], int(k), Dall[gi, k]))
                    found.sort(key=lambda t: (t[0], t[1]))
                    for j, k, D in found:
                        pending[qi].append((k, j, D))
                swept = True
            except (TypeError, ValueError, NotImplementedError, RuntimeError, AttributeError, IndexError):
                g[:] = 0.0  # f is not batch-transparent (same requirement as vmap): one recording per nudge
                pending = [[] for _ in base.queries]
            for k in range(0 if not swept else theta.size, theta.size):
                h = _H_PARAM * max(1.0, abs(flat[k]))
                ap, am = list(args), list(args)
                tp, tm = flat.copy(), flat.copy()
                tp[k] += h
                tm[k] -= h
                ap[i], am[i] = tp.reshape(theta.shape), tm.reshape(theta.shape)
                vp, trp = replay(ap, stored)
                vm, trm = replay(am, stored)
                g[k] += (vp - vm) / (2 * h)  # loss terms that use the parameter outside the circuits
                if len(trp.queries) != len(base.queries) or len(trm.queries) != len(base.queries):
                    raise RuntimeError("value_and_grad: the structure of the function depends on the parameter values")
                for qi, (qp, qm, qb) in enumerate(zip(trp.queries, trm.queries, base.queries)):
                    if len(qp.ops) != len(qb.ops) or len(qm.ops) != len(qb.ops):
                        raise RuntimeError("value_and_grad: the circuit structure depends on the parameter values")
                    if leader[qi] != qi or not any(np.any(dl_de[m]) for m in groups[qb.key]):
                        continue
                    for j, ((_, Mp), (_, Mm)) in enumerate(zip(qp.ops, qm.ops)):
                        D = (Mp - Mm) / (2 * h)
                        if np.abs(D).max() > 1e-12:
                            pending[qi].append((k, j, D))
            for qi, lst in enumerate(pending):
                all_pending[qi].extend((pos, k, j, D) for k, j, D in lst)
        for qi, lst4 in enumerate(all_pending):
            if not lst4:
                continue
            lst = [(k, j, D) for _, k, j, D in lst4]
            members = groups[base.queries[qi].key]
            merged = _Query.__new__(_Query)  # all terms of the group in one batched evaluation
            merged.nqubits, merged.ops = base.queries[qi].nqubits, base.queries[qi].ops
            pauli = [m for m in members if base.queries[m].coo is None]
            merged.fl = [x for m in pauli for x in base.queries[m].fl]
            merged.sg = [x for m in pauli for x in base.queries[m].sg]
            merged.ny = [x for m in pauli for x in base.queries[m].ny]
            wts = np.concatenate([dl_de[m] for m in pauli]) if pauli else np.zeros(0)
            # sparse-operator queries of the group: (operator, d loss / d value)
            merged.coos = [(base.queries[m].coo, float(dl_de[m][0])) for m in members if base.queries[m].coo is not None]
            merged.circ, merged.ny = base.queries[qi].circ, list(merged.ny)
            contrib = _adjoint_contrib(merged, wts, lst, dtype) if _ADJOINT else None
            if contrib is None:
                if merged.coos:
                    raise NotImplementedError("value_and_grad through a sparse-matrix Hamiltonian needs the adjoint sweep: unitary gates, "
                                              "derivative-carrying gates on <= 2 qubits, a Hermitian operator")
                ADJOINT_STATS["fallbacks"] += 1
                diff = _shift_batch(merged, merged, [(j, D) for _, j, D in lst], dtype)  # [nshift, nterms]
                de = 0.5 * np.real(diff)  # d E_t / d theta through that gate
                contrib = de @ wts
            for (pos, k, _, _), c in zip(lst4, contrib):
                gflat[pos][k] += c
        for pos, i in enumerate(nums):
            grads.append(gflat[pos].reshape(np.shape(args[i])))
        gr: Any = grads[0] if single else tuple(grads)
        val: Any = np.asarray(value)
        return ((val, aux), gr) if has_aux else (val, gr)

    return wrapper


def grad(f: Callable[..., Any], argnums: Union[int, Sequence[int]] = 0, has_aux: bool = False) -> Callable[..., Any]:
    vg = value_and_grad(f, argnums=argnums, has_aux=has_aux)

    def wrapper(*args: Any, **kws: Any) -> Any:
        v, g = vg(*args, **kws)
        return (g, v[1]) if has_aux else g

    return wrapper


def vectorized_value_and_grad(f: Callable[..., Any], argnums: Union[int, Sequence[int]] = 0,
                              vectorized_argnums: Union[int, Sequence[int]] = 0, has_aux: bool = False) -> Callable[..., Any]:
    """jax_backend.py:734-776: values per batch element; the gradient is that of the SUM over the
    batch -- so it is per element for an argument that is itself vectorised and summed otherwise."""
    if has_aux:
        raise NotImplementedError("vvag with has_aux")
    single = isinstance(argnums, int)
    nums: Tuple[int, ...] = (argnums,) if single else tuple(argnums)  # type: ignore[assignment]
    vnums: Tuple[int, ...] = (vectorized_argnums,) if isinstance(vectorized_argnums, int) else tuple(vectorized_argnums)
    vg = value_and_grad(f, argnums=nums)

    def wrapper(*args: Any, **kws: Any) -> Any:
        B = int(np.shape(args[vnums[0]])[0])
        vals, per = [], []
        for b in range(B):
            a = [np.asarray(x)[b] if i in vnums else x for i, x in enumerate(args)]
            v, g = vg(*a, **kws)
            vals.append(v)
            per.append(g)
        grads = []
        for pos, i in enumerate(nums):
            gs = np.stack([p[pos] for p in per])
            grads.append(gs if i in vnums else gs.sum(axis=0))
        return np.asarray(vals), (grads[0] if single else tuple(grads))

    return wrapper


# ---- vjp / jvp / jacobians / hessian on top of the sweep (abstract_backend.py:1461-1658) -----------------
def _as_tuple(x: Any) -> Tuple[Tuple[Any, ...], bool]:
    if isinstance(x, (tuple, list)):
        return tuple(x), True
    return (x,), False


def _component(f: Callable[..., Any], i: Optional[int], weights: Optional[np.ndarray] = None) -> Callable[..., Any]:
    """scalar function of the same arguments: component ``i`` of ``f`` (flattened), or sum(weights * f)"""

    def g(*a: Any, **k: Any) -> Any:
        out = f(*a, **k)
        if isinstance(out, (tuple, list)):
            raise NotImplementedError("jacobians / vjp of functions with several outputs")
        flat = out.reshape([-1]) if isinstance(out, BatchArray) else np.reshape(np.asarray(out), [-1])
        if i is not None:
            return flat[i]
        acc = None
        for j, w in enumerate(weights):  # type: ignore[arg-type]
            if w != 0:
                t = flat[j] * w
                acc = t if acc is None else acc + t
        return acc if acc is not None else flat[0] * 0.0

    return g


def vjp(f: Callable[..., Any], inputs: Any, v: Any) -> Tuple[Any, Any]:
    """(f(*inputs), v^T J): the vector-Jacobian product is the gradient of sum(v * f) -- ONE adjoint sweep
    whatever the number of outputs (abstract_backend.py:1484-1507)."""
    ins, many = _as_tuple(inputs)
    value = f(*ins)
    w = np.real(np.asarray(v, dtype=np.complex128)).reshape(-1)
    g = grad(_component(f, None, w), argnums=tuple(range(len(ins))))(*ins)
    return value, (tuple(g) if many else g[0])


def jacrev(f: Callable[..., Any], argnums: Union[int, Sequence[int]] = 0) -> Callable[..., Any]:
    """Jacobian of a vector of expectation-value expressions, output axes first (abstract_backend.py:1574-1648):
    one adjoint sweep per output component."""
    single = isinstance(argnums, int)
    nums: Tuple[int, ...] = (argnums,) if single else tuple(argnums)  # type: ignore[assignment]

    def wrapper(*args: Any, **kws: Any) -> Any:
        values = np.asarray(f(*args, **kws))
        m = int(values.size)
        rows = [grad(_component(f, i), argnums=nums)(*args, **kws) for i in range(m)]
        out = []
        for pos, a in enumerate(nums):
            shape = tuple(values.shape) + tuple(np.shape(args[a]))
            out.append(np.stack([np.asarray(r[pos]) for r in rows]).reshape(shape))
        return out[0] if single else tuple(out)

    return wrapper


jacfwd = jacrev  # same array for one output and one argument; there is no separate forward mode here


def jvp(f: Callable[..., Any], inputs: Any, v: Any) -> Tuple[Any, Any]:
    """(f(*inputs), J v) (abstract_backend.py:1461-1482), assembled from the reverse-mode Jacobian"""
    ins, _ = _as_tuple(inputs)
    vs, _ = _as_tuple(v)
    value = f(*ins)
    J = jacrev(f, argnums=tuple(range(len(ins))))(*ins)
    out_shape = np.shape(np.asarray(value))
    acc = np.zeros(out_shape)
    for Ja, va in zip(J, vs):
        acc = acc + np.tensordot(np.asarray(Ja), np.real(np.asarray(va)), axes=np.ndim(va)).reshape(out_shape)
    return value, acc


def hessian(f: Callable[..., Any], argnums: int = 0, step: float = 1e-4) -> Callable[..., Any]:
    """Hessian of a scalar loss: central differences (step ``step``) of the EXACT adjoint gradient, symmetrised.
    The reference nests forward over reverse mode (abstract_backend.py:1650-1658); there is no second-order sweep here,
    so the result carries a truncation error of O(step^2) and, in complex64, rounding noise of ~1e-7 / step."""
    if not isinstance(argnums, int):
        raise NotImplementedError("hessian with respect to several arguments")
    g = grad(f, argnums=argnums)

    def wrapper(*args: Any, **kws: Any) -> Any:
        x = np.array(np.real(np.asarray(args[argnums])), dtype=np.float64)
        P = x.size
        H = np.zeros((P, P))
        for k in range(P):
            h = step * max(1.0, abs(x.reshape(-1)[k]))
            up, dn = x.copy().reshape(-1), x.copy().reshape(-1)
            up[k] += h
            dn[k] -= h
            a_up, a_dn = list(args), list(args)
            a_up[argnums], a_dn[argnums] = up.reshape(x.shape), dn.reshape(x.shape)
            H[:, k] = (np.asarray(g(*a_up, **kws)).reshape(-1) - np.asarray(g(*a_dn, **kws)).reshape(-1)) / (2 * h)
        H = 0.5 * (H + H.T)
        return H.reshape(x.shape + x.shape)

    return wrapper
