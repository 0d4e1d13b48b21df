"""``Circuit``: the reference's circuit surface on top of the B200 statevector engine.

Mirrors, for the hot path, tensorcircuit/abstractcircuit.py (gate-method registration and
index broadcast, :28-66, :111-326; ``expectation_ps`` :1208-1288; qir helpers),
tensorcircuit/basecircuit.py (``apply_general_gate`` :120-245, ``probability`` :510-523,
``sample`` :525-616) and tensorcircuit/circuit.py (``__init__`` :43-122, ``wavefunction``
:792-812, ``expectation`` :914-990).

Differences that are deliberate: gates are *recorded* (same ``_qir`` dictionaries) and
executed lazily by fused CUDA passes on a device-resident state the first time a query needs
it -- the same moment the reference contracts its network (basecircuit.py:253-257).  After a
query, later gates are applied incrementally to the cached state instead of re-simulating."""

from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import cons, gates
from .batching import BatchArray, batch_of, is_batched
from .fusion import GateOp, fuse, fuse_structured
from .gates import Gate
from .lazy import LazyScalar, TermPool
from .quantum import correlation_from_samples, ps2xyz, sample2all, sample_int2bin

Tensor = Any

sgates = (
    ["i", "x", "y", "z", "h", "t", "s", "td", "sd", "wroot"]
    + ["cnot", "cz", "swap", "cy", "ox", "oy", "oz"]
    + ["toffoli", "fredkin"]
)
vgates = [
    "r", "cr", "u", "cu", "rx", "ry", "rz", "phase", "rxx", "ryy", "rzz", "cphase",
    "crx", "cry", "crz", "orx", "ory", "orz", "iswap", "any", "exp", "exp1",
]
mpogates = ["multicontrol", "mpo"]
gate_aliases = [
    ["cnot", "cx"], ["fredkin", "cswap"], ["toffoli", "ccnot"], ["toffoli", "ccx"],
    ["any", "unitary"], ["sd", "sdg"], ["td", "tdg"],
]


# autodiff.py binds every new engine state to the Circuit that created it through this hook
_STATE_HOOK: Optional[Callable[[Any, Any], None]] = None


def is_sequence(x: Any) -> bool:
    return isinstance(x, (list, tuple, np.ndarray))


class DeviceArray:
    """A tensor that lives on the GPU (the engine's state or a probability vector).

    Converts to numpy on demand (``np.asarray`` / indexing copy device->host), so reference
    test bodies such as ``np.testing.assert_allclose(c.state(), ...)`` read unchanged."""

    def __init__(self, t: Any):
        self.t = t

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.t.shape)

    @property
    def dtype(self) -> Any:
        return np.dtype(str(self.t.dtype).replace("torch.", ""))

    def __len__(self) -> int:
        return self.t.shape[0]

    def reshape(self, *shape: Any) -> "DeviceArray":
        if len(shape) == 1 and is_sequence(shape[0]):
            shape = tuple(shape[0])
        return DeviceArray(self.t.reshape(*shape))

    def numpy(self) -> np.ndarray:
        return self.t.detach().cpu().numpy()

    def torch(self) -> Any:
        return self.t

    def __array__(self, dtype: Any = None, copy: Any = None) -> np.ndarray:
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx: Any) -> Any:
        r = self.t[idx]
        a = r.detach().cpu().numpy()
        return a[()] if a.ndim == 0 else a

    def __repr__(self) -> str:
        return "DeviceArray(shape=%s, dtype=%s, device=%s)" % (self.shape, self.dtype, self.t.device)


def _np_scalar(v: complex, dtype: str) -> np.ndarray:
    return np.array(v, dtype=dtype)


class Circuit:
    """``Circuit`` class: simulate the pure-state evolution of ``nqubits`` qubits."""

    is_dm = False
    sgates = sgates
    vgates = vgates
    mpogates = mpogates
    gate_aliases = gate_aliases
    # Fusion width.  With staged multi-block passes (use_passes) the state is read and written
    # once per *pass*, so narrow blocks are best: they minimise the FP32 work (8*2^k flop per
    # amplitude per block), which is what bounds a pass once several blocks share it.  Without
    # passes k=3 is the widest block that stays HBM-bound (measured).
    fusion_kmax = 2
    use_passes = True
    # structure-aware fusion (fusion.fuse_structured): diagonal / permutation / dense classes
    # kept apart so that the gate pass can treat them differently; dense-only fusion otherwise
    structured_fusion = True

    def _fuse(self, ops: Sequence[GateOp], ntot: int) -> List[Any]:
        if self.structured_fusion and self.use_passes:
            return fuse_structured(ops, ntot, kmax=self.fusion_kmax)
        return fuse(ops, ntot, kmax=self.fusion_kmax)

    def __init__(
        self,
        nqubits: int,
        inputs: Optional[Tensor] = None,
        mps_inputs: Optional[Any] = None,
        split: Optional[Dict[str, Any]] = None,
    ) -> None:
        if mps_inputs is not None:
            raise NotImplementedError("mps_inputs is outside the statevector hot path")
        self._nqubits = int(nqubits)
        self.inputs = inputs
        self.mps_inputs = None
        self.split = split
        self.is_mps = False
        self.circuit_param = {"nqubits": nqubits, "inputs": inputs, "mps_inputs": None, "split": split}
        self._ntot = self._nqubits
        if inputs is not None:
            size = int(np.prod(np.shape(inputs))) if not hasattr(inputs, "numel") else int(inputs.numel())
            n = int(round(np.log(size) / np.log(2)))
            # circuit.py:92 -- 2^n entries, or 2^(2n) (extra trailing legs, "unitary as input")
            assert n == nqubits or n == 2 * nqubits
            assert 2**n == size
            self._ntot = n
        self._qir: List[Dict[str, Any]] = []
        self._extra_qir: List[Dict[str, Any]] = []
        self._ops: List[GateOp] = []
        self._dtype = cons.dtypestr
        self._state = None  # DeviceState
        self._applied = 0
        self._batch: Optional[int] = None
        self._pool: Optional[TermPool] = None  # expectation terms registered on the current state (lazy.py)
        self.state_tensor = None  # kept for source compatibility (basecircuit.py:245)

    # ------------------------------------------------------------------------------------
    # gate application
    # ------------------------------------------------------------------------------------
    def apply_general_gate(
        self,
        gate: Union[Gate, Any],
        *index: int,
        name: Optional[str] = None,
        split: Optional[Dict[str, Any]] = None,
        mpo: bool = False,
        ir_dict: Optional[Dict[str, Any]] = None,
    ) -> None:
        if name is None:
            name = ""
        gate_dict = {"gate": gate, "index": index, "name": name, "split": split, "mpo": mpo}
        if ir_dict is not None:
            ir_dict.update(gate_dict)
        else:
            ir_dict = gate_dict
        self._close_pool()  # pending expectation terms belong to the state before this gate
        self._qir.append(ir_dict)
        assert len(index) == len(set(index))
        index = tuple([i if i >= 0 else self._nqubits + i for i in index])
        if not isinstance(gate, Gate):
            gate = Gate(gate)
        m = gate.matrix()
        d = m.shape[-1]
        if d != 2 ** len(index):
            raise ValueError("gate %s acts on %d legs but %d indices were given" % (name, int(np.log2(d)), len(index)))
        for i in index:
            if not 0 <= i < self._nqubits:
                raise ValueError("qubit index %d out of range" % i)
        if is_batched(m):
            if self._batch is None:
                self._batch = m.batch
                self._state = None  # a batch-1 state cannot be widened in place: re-run
                self._applied = 0
            elif self._batch != m.batch:
                raise ValueError("inconsistent vmap batch sizes")
        self._ops.append(GateOp(index, m, name, getattr(gate, "kind", None) if is_batched(m) else None))
        self.state_tensor = None

    apply = apply_general_gate

    @staticmethod
    def apply_general_variable_gate_delayed(gatef: Callable[..., Gate], name: Optional[str] = None, mpo: bool = False) -> Callable[..., None]:
        if name is None:
            name = getattr(gatef, "n")

        def apply(self: "Circuit", *index: int, **vars: Any) -> None:
            split = None
            localname = name
            if "name" in vars:
                localname = vars.pop("name")
            if "split" in vars:
                split = vars.pop("split")
            gate_dict = {"gatef": gatef, "index": index, "target": list(index), "name": localname, "split": split, "mpo": mpo, "parameters": vars}
            gate = gatef(**vars)
            self.apply_general_gate(gate, *index, name=localname, split=split, mpo=mpo, ir_dict=gate_dict)

        def apply_list(self: "Circuit", *index: Any, **vars: Any) -> None:
            if isinstance(index[0], (int, np.integer)):
                apply(self, *[int(i) for i in index], **vars)
            elif is_sequence(index[0]) or isinstance(index[0], range):
                for i, ind in enumerate(zip(*index)):
                    nvars = {}
                    for k, v in vars.items():
                        try:
                            nvars[k] = v[i]
                        except Exception:  # noqa: BLE001  (abstractcircuit.py:156-162)
                            nvars[k] = v
                    apply(self, *[int(j) for j in ind], **nvars)
            else:
                raise ValueError("Illegal index specification")

        return apply_list

    @staticmethod
    def apply_general_gate_delayed(gatef: Callable[[], Gate], name: Optional[str] = None, mpo: bool = False) -> Callable[..., None]:
        if name is None:
            name = getattr(gatef, "n")
        defaultname = name

        def apply(self: "Circuit", *index: int, split: Optional[Dict[str, Any]] = None, name: Optional[str] = None) -> None:
            localname = name if name is not None else defaultname
            gate = gatef()
            self.apply_general_gate(gate, *index, name=localname, split=split, mpo=mpo, ir_dict={"gatef": gatef})

        def apply_list(self: "Circuit", *index: Any, **kws: Any) -> None:
            if isinstance(index[0], (int, np.integer)):
                apply(self, *[int(i) for i in index], **kws)
            elif is_sequence(index[0]) or isinstance(index[0], range):
                for ind in zip(*index):
                    apply(self, *[int(j) for j in ind], **kws)
            else:
                raise ValueError("Illegal index specification")

        return apply_list

    @classmethod
    def _meta_apply(cls) -> None:
        """Register gate methods by reflection (abstractcircuit.py:217-326)."""
        for g in sgates:
            for nm in (g, g.upper()):
                setattr(cls, nm, cls.apply_general_gate_delayed(gatef=getattr(gates, g), name=g))
                getattr(cls, nm).__doc__ = "Apply **%s** gate on the circuit (gates.%s_gate)." % (g.upper(), g)
        for g in vgates:
            for nm in (g, g.upper()):
                setattr(cls, nm, cls.apply_general_variable_gate_delayed(gatef=getattr(gates, g), name=g))
                getattr(cls, nm).__doc__ = "Apply **%s** gate with parameters on the circuit (gates.%s_gate)." % (g.upper(), g)
        for g in mpogates:
            for nm in (g, g.upper()):
                setattr(cls, nm, cls.apply_general_variable_gate_delayed(gatef=getattr(gates, g), name=g, mpo=True))
        for present, *others in gate_aliases:
            for alias in others:
                setattr(cls, alias, getattr(cls, present))

    # ------------------------------------------------------------------------------------
    # IR helpers (abstractcircuit.py:328-520)
    # ------------------------------------------------------------------------------------
    def to_qir(self) -> List[Dict[str, Any]]:
        return self._qir

    @classmethod
    def from_qir(cls, qir: List[Dict[str, Any]], circuit_params: Optional[Dict[str, Any]] = None) -> "Circuit":
        if circuit_params is None:
            circuit_params = {}
        if "nqubits" not in circuit_params:
            nqubits = max(max(d["index"]) for d in qir) + 1
            circuit_params["nqubits"] = nqubits
        c = cls(**circuit_params)
        c = cls._apply_qir(c, qir)
        return c

    @staticmethod
    def _apply_qir(c: "Circuit", qir: List[Dict[str, Any]]) -> "Circuit":
        for d in qir:
            if "parameters" not in d:
                c.apply_general_gate_delayed(d["gatef"], d["name"], mpo=d["mpo"])(c, *d["index"], split=d["split"])
            else:
                c.apply_general_variable_gate_delayed(d["gatef"], d["name"], mpo=d["mpo"])(c, *d["index"], **d["parameters"], split=d["split"])
        return c

    def append_from_qir(self, qir: List[Dict[str, Any]]) -> None:
        self._apply_qir(self, qir)

    def inverse(self, circuit_params: Optional[Dict[str, Any]] = None) -> "Circuit":
        if circuit_params is None:
            circuit_params = {"nqubits": self._nqubits}
        c = type(self)(**circuit_params)
        for d in reversed(self._qir):
            if "parameters" not in d:
                self.apply_general_gate_delayed(d["gatef"].adjoint(), d["name"], mpo=d["mpo"])(c, *d["index"], split=d["split"])
            else:
                self.apply_general_variable_gate_delayed(d["gatef"].adjoint(), d["name"], mpo=d["mpo"])(c, *d["index"], **d["parameters"], split=d["split"])
        return c

    def append(self, c: "Circuit", indices: Optional[List[int]] = None) -> "Circuit":
        qir = c.to_qir()
        if indices is not None:
            qir_new = []
            for d in qir:
                d = dict(d)
                d["index"] = [indices[i] for i in d["index"]]
                qir_new.append(d)
            qir = qir_new
        self._apply_qir(self, qir)
        return self

    def prepend(self, c: "Circuit") -> "Circuit":
        """abstractcircuit.py:1133-1146: ``c`` runs before the gates recorded so far."""
        newc = type(self).from_qir(c.to_qir() + self.to_qir(), {"nqubits": self._nqubits, "inputs": self.inputs})
        self.__dict__.update(newc.__dict__)
        return self

    def copy(self) -> "Circuit":
        """abstractcircuit.py:1192-1195: a new circuit with the same record (and its own state)."""
        return type(self).from_qir(self.to_qir(), {"nqubits": self._nqubits, "inputs": self.inputs})

    def gate_count_by_condition(self, cond_func: Callable[[Dict[str, Any]], bool]) -> int:
        """abstractcircuit.py:655-681"""
        return sum(1 for d in self._qir if cond_func(d))

    def _instruction(self, name: str, index: Any) -> None:
        self._extra_qir.append({"index": index, "name": name, "gatef": name, "instruction": True, "pos": len(self._qir)})

    def measure_instruction(self, *index: int) -> None:
        """abstractcircuit.py:695-711: a flag for translators, no effect on the simulation."""
        for ind in index:
            self._instruction("measure", [ind])

    def reset_instruction(self, *index: int) -> None:
        """abstractcircuit.py:713-729"""
        for ind in index:
            self._instruction("reset", [ind])

    def barrier_instruction(self, *index: Any) -> None:
        """abstractcircuit.py:731-746"""
        self._instruction("barrier", index)

    def is_valid(self) -> bool:
        """circuit.py:776-790 checks the wiring of the node graph; there is no graph here, the
        record is valid by construction (arity and range are checked when a gate is applied)."""
        return True

    def gate_count(self, gate_list: Optional[Union[str, Sequence[str]]] = None) -> int:
        if gate_list is None:
            return len(self._qir)
        if isinstance(gate_list, str):
            gate_list = [gate_list]
        return sum(1 for d in self._qir if d["name"] in gate_list)

    def gate_summary(self) -> Dict[str, int]:
        s: Dict[str, int] = {}
        for d in self._qir:
            s[d["name"]] = s.get(d["name"], 0) + 1
        return s

    def _close_pool(self) -> None:
        """Evaluate the expectation terms registered so far (one launch group) and start afresh:
        called whenever the state is about to change."""
        if self._pool is not None:
            self._pool.flush()
            self._pool.closed = True
            self._pool = None

    # Deferred evaluation of expectation_ps on unbatched circuits (lazy.py): a Python loop over
    # the terms of a Hamiltonian costs one launch group instead of one launch + sync per term.
    lazy_expectation = True

    def replace_inputs(self, inputs: Tensor) -> None:
        """basecircuit.py:805-822: same circuit, new initial state."""
        self._close_pool()
        self.inputs = inputs
        self._state = None
        self._applied = 0

    # ------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------
    def _ensure_state(self) -> Any:
        from .engine import DeviceState

        batch = self._batch or 1
        if self._state is None or self._state.batch != batch:
            if cons.distributed_state:
                from .parallel import world

                if world()[1] > 1:
                    from .dist import DistEngineState as DeviceState  # noqa: F811
            st = DeviceState(self._ntot, self._dtype, batch)
            if self.inputs is None:
                st.init_zero()
            else:
                src = self.inputs.t if isinstance(self.inputs, DeviceArray) else self.inputs
                st.load(src)
            self._state = st
            self._applied = 0
            if _STATE_HOOK is not None:
                _STATE_HOOK(self, st)
        if self._applied < len(self._ops):
            pending = self._ops[self._applied :]
            blocks = self._fuse(pending, self._ntot)
            if self.use_passes and hasattr(self._state, "apply_planned"):
                self._state.apply_planned(blocks)
            else:
                self._state.apply_blocks(blocks)
            self._applied = len(self._ops)
        return self._state

    def wavefunction(self, form: str = "default") -> Any:
        """circuit.py:792-810.  Returns a :class:`DeviceArray` (a copy below 2^30 amplitudes, a
        zero-copy view of the live state above, where a copy would not fit next to it)."""
        st = self._ensure_state()
        t = st.buf
        if self._batch is None:
            t = t[0]
            if self._ntot < 30:
                t = t.clone()
            shape = {"default": [-1], "ket": [-1, 1], "bra": [1, -1]}[form]
            return DeviceArray(t.reshape(shape))
        return BatchedDeviceArray(t.clone())

    state = wavefunction

    def matrix(self) -> np.ndarray:
        """Unitary of the circuit: evolve the identity (circuit.py:814-826 equivalent)."""
        n = self._nqubits
        c = type(self)(n, inputs=np.eye(2**n))
        self._apply_qir(c, self._qir)
        return np.asarray(c.wavefunction()).reshape(2**n, 2**n)

    def amplitude(self, l: Union[str, Sequence[int]]) -> Any:
        if isinstance(l, str):
            bits = [int(s) for s in l]
        else:
            bits = [int(b) for b in l]
        assert len(bits) == self._nqubits
        idx = 0
        for b in bits:
            idx = (idx << 1) | b
        st = self._ensure_state()
        if self._ntot != self._nqubits:
            raise NotImplementedError("amplitude with unitary-form inputs")
        return DeviceArray(st.buf[0])[idx]

    def probability(self) -> Any:
        """basecircuit.py:510-523"""
        st = self._ensure_state()
        p = st.probability()
        return DeviceArray(p[0]) if self._batch is None else BatchedDeviceArray(p)

    def _bitpos(self, q: int) -> int:
        return self._ntot - 1 - q

    def _pauli_masks(self, x: Sequence[int], y: Sequence[int], z: Sequence[int]) -> Tuple[int, int, int]:
        occupied = set()
        fl = sg = 0
        for lst, isx, isz in ((x, True, False), (y, True, True), (z, False, True)):
            for i in lst:
                i = int(i)
                i = i if i >= 0 else self._nqubits + i
                if i in occupied:
                    raise ValueError("Cannot measure two operators in one index")  # basecircuit.py:306
                occupied.add(i)
                m = 1 << self._bitpos(i)
                if isx:
                    fl |= m
                if isz:
                    sg |= m
        return fl, sg, len(y)

    def expectation_ps_many(self, pss: Sequence[Sequence[int]]) -> Any:
        """All Pauli strings ``pss`` ([nterms][nqubits] of 0..3) in as few reads of the state as
        their flip masks allow (one for a tile-local Hamiltonian).  Returns complex [nterms]."""
        st = self._ensure_state()
        fl, sg, ny = [], [], []
        for ps in pss:
            d = ps2xyz(list(ps))
            f, s, n = self._pauli_masks(d["x"], d["y"], d["z"])
            fl.append(f)
            sg.append(s)
            ny.append(n)
        r = st.expectation_terms(fl, sg, ny)
        if self._batch is None:
            return r[0].astype(self._dtype)
        return BatchArray(r.astype(self._dtype))

    def expectation_ps(
        self,
        x: Optional[Sequence[int]] = None,
        y: Optional[Sequence[int]] = None,
        z: Optional[Sequence[int]] = None,
        ps: Optional[Sequence[int]] = None,
        reuse: bool = True,
        noise_conf: Optional[Any] = None,
        nmc: int = 1000,
        status: Optional[Tensor] = None,
        **kws: Any,
    ) -> Tensor:
        """abstractcircuit.py:1208-1288 (``ps`` overrides x/y/z)."""
        if noise_conf is not None:
            raise NotImplementedError("noise_conf is outside the statevector hot path")
        if ps is not None:
            d = ps2xyz(list(ps))
            x, y, z = d.get("x"), d.get("y"), d.get("z")
        fl, sg, ny = self._pauli_masks(x or [], y or [], z or [])
        st = self._ensure_state()  # (a state hook may turn the circuit into a batched one: autodiff replays)
        if self._batch is None and self.lazy_expectation:
            if self._pool is None:
                self._pool = TermPool(self)
            return LazyScalar([(self._pool, self._pool.add(fl, sg, ny), 1.0 + 0.0j)], 0j, "c", self._dtype)
        r = st.expectation_terms([fl], [sg], [ny])
        if self._batch is None:
            return _np_scalar(r[0, 0], self._dtype)
        return BatchArray(r[:, 0].astype(self._dtype))

    def expectation(self, *ops: Tuple[Any, List[int]], reuse: bool = True, enable_lightcone: bool = False,
                    noise_conf: Optional[Any] = None, nmc: int = 1000, status: Optional[Tensor] = None, **kws: Any) -> Tensor:
        """circuit.py:914-990 for operators on disjoint sites: each operator is expanded in the
        Pauli basis and the resulting strings go through the multi-term expectation kernel."""
        if noise_conf is not None:
            raise NotImplementedError("noise_conf is outside the statevector hot path")
        occupied = set()
        # list of (coefficient, {qubit: pauli}) partial strings
        terms: List[Tuple[Any, Dict[int, int]]] = [(1.0 + 0.0j, {})]
        for op, index in ops:
            if isinstance(index, (int, np.integer)):
                index = [int(index)]
            index = [int(i) if i >= 0 else self._nqubits + int(i) for i in index]
            for e in index:
                if e in occupied:
                    raise ValueError("Cannot measure two operators in one index")
                occupied.add(e)
            m = op.matrix() if isinstance(op, Gate) else gates.reshapem(op.tensor if hasattr(op, "tensor") else op)
            dec = _pauli_decompose(m, len(index))
            new_terms = []
            for c0, d0 in terms:
                for c1, pl in dec:
                    d = dict(d0)
                    for q, p in zip(index, pl):
                        if p:
                            d[q] = p
                    new_terms.append((c0 * c1, d))
            terms = new_terms
            if len(terms) > 4096:
                raise NotImplementedError("operator product expands into too many Pauli strings")
        pss = []
        for _, d in terms:
            ps = [0] * self._nqubits
            for q, p in d.items():
                ps[q] = p
            pss.append(ps)
        st = self._ensure_state()  # (a state hook may turn the circuit into a batched one: autodiff replays)
        if self._batch is None and self.lazy_expectation:
            # the other reference idiom for an energy -- a loop over c.expectation((gates.x(), [i])) ... ,
            # benchmarks/scripts/vqe_tc.py:74-81 -- joins the same pool of pending terms as expectation_ps: all
            # calls on one state are evaluated as one launch group when a value is read
            if self._pool is None:
                self._pool = TermPool(self)
            lterms = []
            for (c, _), ps in zip(terms, pss):
                d = ps2xyz(ps)
                fl, sg, ny = self._pauli_masks(d["x"], d["y"], d["z"])
                lterms.append((self._pool, self._pool.add(fl, sg, ny), complex(c)))
            return LazyScalar(lterms, 0j, "c", self._dtype)
        vals = self.expectation_ps_many(pss)
        tot: Any = 0.0
        for i, (c, _) in enumerate(terms):
            tot = tot + c * vals[i]
        if is_batched(tot):
            return tot.astype(self._dtype)
        return _np_scalar(tot, self._dtype)

    # ------------------------------------------------------------------------------------
    # sampling (basecircuit.py:525-616)
    # ------------------------------------------------------------------------------------
    def sample(
        self,
        batch: Optional[int] = None,
        allow_state: bool = False,
        readout_error: Optional[Sequence[Any]] = None,
        format: Optional[str] = None,
        random_generator: Optional[Any] = None,
        status: Optional[Tensor] = None,
        format_: Optional[str] = None,
    ) -> Any:
        """Sample bitstrings from the final state with the reference's CDF rule
        (abstract_backend.py:1145-1157): ``r = total*(1-u)``, first index with CDF >= r.

        ``allow_state=False`` (the reference's qubit-by-qubit branch, basecircuit.py:558-585)
        draws from the same exact distribution through the same state sampler here."""
        if format_ is not None:
            format = format_
        if self._batch is not None:
            # vmap over ``status`` and / or circuit parameters (docs/source/advance.rst:95-170): one sampler
            # run per batch row, on a view of that row of the batched state
            if format not in ("sample_int", "sample_bin"):
                raise NotImplementedError("sample() inside vmap returns arrays: use format='sample_int' or 'sample_bin'")
            outs = [rc.sample(batch=batch, allow_state=allow_state, readout_error=readout_error, format=format, random_generator=random_generator,
                              status=sb) for rc, sb in self._row_circuits(status)]
            return BatchArray(np.stack([np.asarray(o) for o in outs]))
        nbatch = 1 if batch is None else int(batch)
        if status is None:
            u = cons.backend.stateful_randu(random_generator, shape=[nbatch]) if random_generator is not None else cons.backend.implicit_randu(shape=[nbatch])
        elif hasattr(status, "is_pinned"):  # torch tensor (e.g. pinned host memory): no host copy
            u = status
            if u.numel() != nbatch:
                raise ValueError("status must hold one uniform per shot")
        else:
            u = np.asarray(status, dtype=np.float64).reshape(-1)
            if u.shape[0] != nbatch:
                raise ValueError("status must hold one uniform per shot")
        st = self._ensure_state()
        if self._ntot != self._nqubits:
            raise NotImplementedError("sample with unitary-form inputs")
        if readout_error is not None:  # basecircuit.py:592-596: sample the readout-corrupted distribution
            st = self._readout_state(readout_error)
        ch = st.sample(u)
        if format is None:  # backward-compatible tuple form (basecircuit.py:609-615)
            import torch

            confg = sample_int2bin(ch, self._nqubits)
            amp = st.buf[0][torch.from_numpy(ch).to(st.buf.device)].cpu().numpy()
            prob = (np.abs(amp) ** 2).astype(cons.rdtypestr)
            r = list(zip(confg, prob))
            return r[0] if batch is None else r
        return sample2all(sample=ch, n=self._nqubits, format=format, jittable=True)

    def _row_circuits(self, status: Any = None) -> List[Tuple["Circuit", Any]]:
        """Unbatched views of a vmapped circuit: (circuit whose state IS row b of the batched state, row b of
        ``status`` when that is batched too)."""
        st = self._ensure_state()
        if not hasattr(st, "row_state"):
            raise NotImplementedError("per-row queries on this kind of state")
        rows = []
        for b in range(st.batch):
            rc = type(self)(self._nqubits)
            rc._ntot = self._ntot
            rc._state = st.row_state(b)
            rc._applied = 0
            sb = status
            if is_batched(status):
                sb = np.asarray(status.a)[b]
            rows.append((rc, sb))
        return rows

    def _readout_state(self, readout_error: Sequence[Any]) -> Any:
        """A state-shaped device buffer whose |amplitude|^2 is the distribution after readout error
        (basecircuit.py:760-803: p' = (tensor product of [[p0|0, 1-p1|1], [1-p0|0, p1|1]]) p).  The
        probabilities sit in the real parts of a complex buffer, the per-qubit stochastic matrices
        go through the ordinary fused passes, then a square root makes it samplable."""
        st = self._ensure_state()
        if self._ntot != self._nqubits or not hasattr(st, "probability_state"):
            raise NotImplementedError("readout_error needs a single-device pure state")
        if len(readout_error) != self._nqubits:
            raise ValueError("readout_error needs one [p0|0, p1|1] pair per qubit")
        r = st.probability_state()
        ops = []
        for q, e in enumerate(readout_error):
            p0, p1 = float(np.real(e[0])), float(np.real(e[1]))
            ops.append(GateOp((q,), np.array([[p0, 1 - p1], [1 - p0, p1]], dtype=np.complex128), "readout"))
        r.apply_planned(self._fuse(ops, self._nqubits)) if hasattr(r, "apply_planned") else r.apply_blocks(self._fuse(ops, self._nqubits))
        r.sqrt_real_inplace()
        return r

    def readouterror_bs(self, readout_error: Optional[Sequence[Any]] = None, p: Optional[Any] = None) -> Any:
        """basecircuit.py:760-803: the bit-string probabilities after readout error, as a device
        array.  ``p`` (the reference's second argument) is always the circuit's own distribution here."""
        if p is not None:
            raise NotImplementedError("readouterror_bs acts on the circuit's own probabilities")
        if readout_error is None:
            return self.probability()
        r = self._readout_state(readout_error)
        return DeviceArray(r.probability()[0])

    def sample_expectation_ps(
        self,
        x: Optional[Sequence[int]] = None,
        y: Optional[Sequence[int]] = None,
        z: Optional[Sequence[int]] = None,
        shots: Optional[int] = None,
        random_generator: Optional[Any] = None,
        status: Optional[Tensor] = None,
        readout_error: Optional[Sequence[Any]] = None,
        noise_conf: Optional[Any] = None,
        **kws: Any,
    ) -> Tensor:
        """Measurement-based Pauli-string expectation (basecircuit.py:618-758): rotate the x sites
        with H and the y sites with rx(pi/2), then the string is a product of Z's -- evaluated
        exactly from the probabilities (``shots=None``: one pass of the diagonal-string kernel)
        or from ``shots`` samples drawn with ``status``.

        The reference rotates a copy of the state; a copy does not fit at 34 qubits, so the
        rotation is applied in place and undone afterwards (the recorded circuit is unchanged)."""
        if noise_conf is not None:
            raise NotImplementedError("noise_conf is outside the statevector hot path")
        x, y, z = list(x or []), list(y or []), list(z or [])
        for i in x + y + z:  # validate before anything is mutated
            if not -self._nqubits <= int(i) < self._nqubits:
                raise ValueError("qubit index %d out of range" % i)
        if shots is None and readout_error is None:
            # exact value: the rotated Z string IS the original Pauli string -- no rotation, no
            # mutation of the cached state, and the query stays differentiable
            r = self.expectation_ps(x=x, y=y, z=z)
            return r.real if is_batched(r) else np.real(r)
        n_ops, n_qir = len(self._ops), len(self._qir)
        self._ensure_state()
        done: List[Tuple[str, int]] = []  # rotations that actually reached the record
        try:
            for i in x:
                self.h(i)
                done.append(("h", i))
            for i in y:
                self.rx(i, theta=np.pi / 2)
                done.append(("rx", i))
            if shots is None:  # sum_e p'_e (-1)^{bits}: the diagonal string on the noisy distribution
                _, sign, _ = self._pauli_masks([], [], x + y + z)
                rs = self._readout_state(readout_error)
                r = float(np.real(rs.expectation_terms([0], [sign], [0])[0, 0]))
            else:
                s = self.sample(batch=int(shots), allow_state=True, readout_error=readout_error, random_generator=random_generator,
                                status=status, format="sample_bin")
                r = correlation_from_samples(x + y + z, np.asarray(s), self._nqubits)
        finally:
            try:
                # undo exactly the rotations that were applied to the device state
                applied = self._applied - n_ops  # gates of `done` already executed on the device
                for name, i in reversed(done[:max(0, applied)]):
                    if name == "h":
                        self.h(i)
                    else:
                        self.rx(i, theta=-np.pi / 2)
                if applied > 0:
                    # run only the inverses: the not-yet-applied forward rotations are dropped
                    pending = self._ops[n_ops + len(done):]
                    st = self._state
                    blocks = self._fuse(pending, self._ntot)
                    if self.use_passes and hasattr(st, "apply_planned"):
                        st.apply_planned(blocks)
                    else:
                        st.apply_blocks(blocks)
            finally:
                del self._ops[n_ops:]
                del self._qir[n_qir:]
                self._applied = n_ops
        return r

    sexpps = sample_expectation_ps

    def measure(self, *index: int, with_prob: bool = False, status: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """Qubit-by-qubit measurement of the final state with the reference's rule
        (basecircuit.py:365-443): for the k-th measured qubit j, ``pu`` = probability of outcome 0
        given the outcomes so far (the reduced-density element the reference contracts; here one
        masked-norm pass over the state), outcome = ``sign(status[k] - pu + 0.31415926e-12)/2 + 0.5``,
        and the running probability is updated as ``p * (pu * (-1)**outcome + outcome)``.
        Returns (outcomes as a real vector, probability of the record or -1.0)."""
        if self._batch is not None:
            outs = [rc.measure(*index, with_prob=with_prob, status=sb) for rc, sb in self._row_circuits(status)]
            return BatchArray(np.stack([o[0] for o in outs])), BatchArray(np.asarray([o[1] for o in outs]))
        st = self._ensure_state()
        if self._ntot != self._nqubits:
            raise NotImplementedError("measure with unitary-form inputs")
        rd = np.float32 if self._dtype == "complex64" else np.float64
        idx = [i if i >= 0 else self._nqubits + i for i in index]
        if status is not None:
            status = np.asarray(status).reshape(-1)
            if status.shape[0] < len(idx):
                raise ValueError("status must hold one uniform per measured qubit")
        eps = 0.31415926 * 1e-12
        sample: List[Any] = []
        p = rd(1.0)
        mask = value = 0
        for k, j in enumerate(idx):
            bit = 1 << self._bitpos(j)
            pu = rd(st.masked_norm2(mask | bit, value) / float(p))
            r = cons.backend.implicit_randu()[0] if status is None else status[k]
            r = rd(np.real(r))
            sign = rd(np.sign(r - pu + eps) / 2 + 0.5)  # 0.5 only if status sits exactly on pu - eps
            sample.append(sign)
            p = rd(p * (pu * rd(-1.0) ** sign + sign)) if sign in (0.0, 1.0) else rd(p * 0.5)
            mask |= bit
            if sign > 0.5:
                value |= bit
        out = np.asarray(sample, dtype=rd)
        return (out, p) if with_prob else (out, -1.0)

    def perfect_sampling(self, status: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """basecircuit.py:359-363: one bitstring by measuring every qubit in order, with its probability."""
        return self.measure(*range(self._nqubits), with_prob=True, status=status)

    measure_jit = measure

    # ------------------------------------------------------------------------------------
    # Monte-Carlo noise trajectories driven by ``status`` (circuit.py:302-352, 447-744,
    # basecircuit.py:824-857).  Everything O(2^n) goes through the same kernels: a Kraus branch
    # is one more (possibly non-unitary, possibly vmap-batched) gate; branch probabilities of a
    # general channel are expectations of K^dagger K on the channel's qubits.
    # ------------------------------------------------------------------------------------
    def mid_measurement(self, index: int, keep: int = 0) -> Tensor:
        """Post-selection on |keep> of qubit ``index``; the state is NOT renormalised
        (circuit.py:302-347)."""
        proj = np.zeros((2, 2), dtype=np.complex128)
        k = 0 if keep < 0.5 else 1
        proj[k, k] = 1.0
        self.any(index, unitary=proj, name="post-select")
        return np.asarray(keep).astype("int32")

    mid_measure = mid_measurement
    post_select = mid_measurement
    post_selection = mid_measurement

    def unitary_kraus(self, kraus: Sequence[Any], *index: int, prob: Optional[Sequence[float]] = None,
                      status: Optional[float] = None, name: Optional[str] = None) -> Tensor:
        """Apply one of ``kraus`` chosen by ``status`` against the cumulative ``prob``
        (circuit.py:473-565).  ``status`` / ``prob`` may be vmap-batched."""
        mats = [gates.reshapem(k.tensor if isinstance(k, Gate) else k) for k in kraus]
        mats = [m if is_batched(m) else np.asarray(m, dtype=np.complex128) for m in mats]
        if prob is None:
            prob = []
            for m in mats:
                a = m.a if is_batched(m) else m
                w = np.real(np.einsum("...ij,...ij->...", a.conj(), a)) / a.shape[-1]
                prob.append(BatchArray(w) if is_batched(m) else w)
            with np.errstate(divide="ignore", invalid="ignore"):
                mats = [
                    BatchArray(m.a / np.sqrt(p.a + 0j)[:, None, None]) if is_batched(m) else m / np.sqrt(p + 0j)
                    for m, p in zip(mats, prob)
                ]
        l = len(mats)
        if status is None:
            status = cons.backend.implicit_randu()[0]
        status = np.real(status) if not is_batched(status) else status.real
        B = batch_of(status, *prob, *mats)
        # cumulative probabilities; the step function of circuit.py:538-548:
        #   r = int( sum_{i<l-1} sign(status - cum_i) / 2 + (l-1)/2 )
        def raw(x: Any) -> np.ndarray:
            return x.a if is_batched(x) else np.asarray(x)

        pr = [np.real(raw(p)).astype(np.float64) for p in prob]
        if B is not None:
            pr = [np.broadcast_to(p, (B,)) for p in pr]
        cum = np.cumsum(np.stack(pr, axis=-1), axis=-1)  # [l] or [B, l]
        st = np.asarray(raw(status), dtype=np.float64)
        if l == 1:
            r = np.zeros(st.shape, dtype=np.int32)
        else:
            r = (np.sum(np.sign(st[..., None] - cum[..., : l - 1]), axis=-1) / 2.0 + (l - 1) / 2.0).astype(np.int32)
        if B is None:
            g = mats[int(r)]
            self.any(*index, unitary=g, name=name if name is not None else "unitary_kraus")
            return np.asarray(r)
        stack = np.stack([np.broadcast_to(raw(m), (B,) + tuple(raw(m).shape[-2:])) for m in mats], axis=1)  # [B, l, d, d]
        rb = np.broadcast_to(r, (B,))
        g = stack[np.arange(B), rb]
        self.any(*index, unitary=BatchArray(g), name=name if name is not None else "unitary_kraus")
        return BatchArray(rb.astype(np.int32))

    unitary_kraus2 = unitary_kraus

    def general_kraus(self, kraus: Sequence[Any], *index: int, status: Optional[float] = None,
                      with_prob: bool = False, name: Optional[str] = None) -> Tensor:
        """Monte-Carlo trajectory step of a general Kraus channel (circuit.py:635-723): branch
        probabilities p_i = <psi| K_i^dagger K_i |psi> on the current state, then the branch chosen
        by ``status`` is applied as K_i / (sqrt(p_i) + 1e-10)."""
        mats = [np.asarray(gates.reshapem(k.tensor if isinstance(k, Gate) else k), dtype=np.complex128) for k in kraus]
        prob = []
        for m in mats:
            kk = m.conj().T @ m
            prob.append(np.real(self.expectation((kk, list(index)))))
        eps = 1e-10
        new = []
        for m, w in zip(mats, prob):
            if is_batched(w):
                new.append(BatchArray(m[None, :, :] / (np.sqrt(w.a)[:, None, None] + eps)))
            else:
                new.append(m / (np.sqrt(w) + eps))
        pick = self.unitary_kraus(new, *index, prob=prob, status=status, name=name)
        return (pick, prob) if with_prob else pick

    apply_general_kraus = general_kraus

    def cond_measurement(self, index: int, status: Optional[float] = None) -> Tensor:
        """Z-basis measurement with collapse, returning the outcome (basecircuit.py:824-857)."""
        return self.general_kraus([np.array([[1.0, 0], [0, 0]]), np.array([[0, 0], [0, 1.0]])], index, status=status, name="measure")

    cond_measure = cond_measurement

    def conditional_gate(self, which: Tensor, kraus: Sequence[Any], *index: int) -> None:
        """Apply ``kraus[which]`` (abstractcircuit.py conditional_gate); ``which`` may be batched."""
        mats = [np.asarray(gates.reshapem(k.tensor if isinstance(k, Gate) else k), dtype=np.complex128) for k in kraus]
        if is_batched(which):
            sel = np.stack(mats)[which.a.astype(np.int64)]
            self.any(*index, unitary=BatchArray(sel), name="conditional")
        else:
            self.any(*index, unitary=mats[int(which)], name="conditional")

    select_gate = conditional_gate

    def depolarizing2(self, index: int, *, px: float, py: float, pz: float, status: Optional[float] = None) -> float:
        """circuit.py:354-386: x / y / z / i chosen by status against px, px+py, px+py+pz."""
        ks = [gates._x_matrix, gates._y_matrix, gates._z_matrix, gates._i_matrix]
        self.unitary_kraus(ks, index, prob=[px, py, pz, 1 - px - py - pz], status=status)
        return 0.0

    @staticmethod
    def apply_general_kraus_delayed(krausf: Callable[..., Sequence[Gate]], is_unitary: bool = False) -> Callable[..., None]:
        def apply(self: "Circuit", *index: int, status: Optional[float] = None, name: Optional[str] = None, **vars: float) -> None:
            kraus = krausf(**vars)
            if not is_unitary:
                self.apply_general_kraus(kraus, *index, status=status, name=name)
            else:
                self.unitary_kraus(kraus, *index, status=status, name=name)

        return apply

    @classmethod
    def _meta_apply_channels(cls) -> None:
        from . import channels as _ch

        for k in _ch.channels:
            setattr(cls, k, cls.apply_general_kraus_delayed(getattr(_ch, k + "channel"), is_unitary=k in ("depolarizing", "generaldepolarizing")))


class BatchedDeviceArray(DeviceArray):
    """vmap result that stays on the device: leading axis is the batch axis."""

    batched_on_device = True


_PAULIS = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])]


def _pauli_decompose(m: Any, k: int) -> List[Tuple[Any, Tuple[int, ...]]]:
    """O = sum_P c_P P over k-qubit Pauli strings; c_P = Tr(P O)/2^k.  Batched O gives batched c."""
    out: List[Tuple[Any, Tuple[int, ...]]] = []
    raw = m.a if is_batched(m) else np.asarray(m, dtype=np.complex128)
    d = 2**k
    for code in range(4**k):
        pl = tuple((code >> (2 * (k - 1 - j))) & 3 for j in range(k))
        p = np.array([[1.0]], dtype=np.complex128)
        for a in pl:
            p = np.kron(p, _PAULIS[a])
        c = np.einsum("ij,...ji->...", p, raw) / d
        if is_batched(m):
            if np.any(np.abs(c) > 1e-15):
                out.append((BatchArray(c), pl))
        elif abs(c) > 1e-15:
            out.append((complex(c), pl))
    return out


Circuit._meta_apply()
Circuit._meta_apply_channels()


def expectation(*ops: Tuple[Any, List[int]], ket: Tensor, bra: Optional[Tensor] = None, conj: bool = True, normalization: bool = False) -> Tensor:
    """tensorcircuit/circuit.py:997-1129: <bra| ops |ket>.  With a separate ``bra`` the operators are
    applied to a copy of ``ket`` as (generally non-unitary) gates and the inner product with ``bra`` is
    one launch of the transition-element kernel; ``conj=False`` gives the bilinear form sum_r bra_r (ops ket)_r."""
    size = int(np.prod(np.shape(ket))) if not hasattr(ket, "numel") else int(ket.numel())
    n = int(round(np.log2(size)))
    c = Circuit(n, inputs=ket)
    if bra is None and conj:
        r = c.expectation(*ops)
        if normalization:
            r = r / c._ensure_state().norm2()[0]
        return r
    nk = float(c._ensure_state().norm2()[0]) if normalization else 1.0
    occupied: set = set()
    for op, index in ops:
        index = [index] if isinstance(index, int) else list(index)
        for e in index:
            if e in occupied:
                raise ValueError("Cannot measure two operators in one index")
            occupied.add(e)
        c.any(*index, unitary=op)
    if bra is None:
        bra = ket
    if not conj:
        bra = bra.t.conj().resolve_conj() if isinstance(bra, DeviceArray) else np.conj(np.asarray(bra))
    cb = Circuit(n, inputs=bra)
    sb = cb._ensure_state()
    v = c._ensure_state().inner(sb)
    if normalization:
        v = v / np.sqrt(nk * float(sb.norm2()[0]))
    return _np_scalar(v, c._dtype)
