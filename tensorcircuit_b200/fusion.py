"""Gate fusion: recorded gates -> dense blocks of <= kmax qubits (north-star item 1).

This is the B200 counterpart of what the reference does at graph level before contraction
(``_merge_single_gates``, tensorcircuit/cons.py:236-279, absorbs rank-2 nodes into a
neighbour): consecutive gates are merged greedily into 2^k x 2^k unitaries so that the state
is read and written once per block instead of once per gate.

* matrices are multiplied on the host in complex128 (and cast once, at upload);
* the grouping depends only on the circuit *structure* (which qubits each gate touches), so
  it is cached by structure -- a VQE loop that rebuilds the same circuit with new angles every
  step only redoes the small matrix products (this is what ``backend.jit`` amounts to here);
* the width cap comes from a roofline cost model: a dense block costs 8*2^k flop per
  amplitude against 16 B of HBM traffic (complex64), so k=4 is the widest block that stays
  HBM-bound on B200 CUDA cores; k=5 is only taken when it removes a whole pass.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .batching import BatchArray, is_batched


MAX_BLOCK_K = 5  # TCB200_MAX_K: widest dense block the kernels apply


@dataclass
class GateOp:
    """One recorded gate: ``qubits`` in the caller's order, ``matrix`` [D, D] or batched."""

    qubits: Tuple[int, ...]
    matrix: Any  # np.ndarray [D, D] complex128, or BatchArray with shape [D, D]
    name: str = ""
    kind: Optional[str] = None  # zero-pattern class when the gate constructor knows it (vmap fast path)

    @property
    def batched(self) -> bool:
        return is_batched(self.matrix)


@dataclass
class Block:
    """A fused block.  ``qubits`` ascending (TC numbering); ``bits`` = amplitude-index bit
    positions ascending; ``matrix`` is [D, D] (or [B, D, D]) with index bit j <-> bits[j]."""

    qubits: Tuple[int, ...]
    bits: Tuple[int, ...]
    matrix: np.ndarray
    batched: bool
    ngates: int
    diagonal: bool = False
    kind: str = "dense"  # 'dense' | 'diag' | 'perm' | 'mono' (fuse_structured)


def embed_apply(block_m: np.ndarray, block_qubits: Sequence[int], g: np.ndarray, gq: Sequence[int]) -> np.ndarray:
    """block_m <- (g on qubits gq) @ block_m, where block_m's row/column index is big-endian in
    the ascending qubit list ``block_qubits`` (= little-endian in ascending bit position).
    Leading batch axes of either operand broadcast."""
    k = len(block_qubits)
    kg = len(gq)
    D = 1 << k
    pos = [block_qubits.index(q) for q in gq]
    batched = block_m.ndim == 3 or g.ndim == 3
    if not batched and k <= 3:
        # small blocks: embed the gate by fancy indexing and multiply (a few microseconds; the
        # general tensordot path below costs ~60)
        og, same_rest = _embed_index(k, tuple(pos))
        return np.matmul(g[og[:, None], og[None, :]] * same_rest, block_m)
    if not batched:
        t = block_m.reshape([2] * k + [D])
        gt = g.reshape([2] * (2 * kg))
        t = np.tensordot(gt, t, axes=(list(range(kg, 2 * kg)), pos))
        t = np.moveaxis(t, list(range(kg)), pos)
        return np.ascontiguousarray(t).reshape(D, D)
    # batched (vmap): embed the gate into the block's index space by fancy indexing and use
    # one batched matmul -- an order of magnitude cheaper on the host than an einsum per gate
    og, same_rest = _embed_index(k, tuple(pos))
    full = g[..., og[:, None], og[None, :]] * same_rest
    return np.matmul(full, block_m)


_EMBED_CACHE: Dict[Any, Tuple[np.ndarray, np.ndarray]] = {}


def _embed_index(k: int, pos: Tuple[int, ...]) -> Tuple[np.ndarray, np.ndarray]:
    """For a block of k qubits (big-endian index over its ascending qubit list) and a gate on the
    block positions ``pos`` (gate order): the gate index of every block index, and the mask of
    index pairs that agree on all other positions."""
    hit = _EMBED_CACHE.get((k, pos))
    if hit is not None:
        return hit
    D = 1 << k
    kg = len(pos)
    idx = np.arange(D)
    og = np.zeros(D, dtype=np.int64)
    rest = idx.copy()
    for a, p in enumerate(pos):
        bit = (idx >> (k - 1 - p)) & 1
        og |= bit << (kg - 1 - a)
        rest &= ~(1 << (k - 1 - p))
    same = (rest[:, None] == rest[None, :]).astype(np.complex128)
    _EMBED_CACHE[(k, pos)] = (og, same)
    return og, same


def _raw(m: Any) -> np.ndarray:
    return m.a if is_batched(m) else np.asarray(m)


class FusionPlan:
    """Grouping of gate indices into blocks for one circuit structure."""

    def __init__(self, groups: List[List[int]], block_qubits: List[Tuple[int, ...]]):
        self.groups = groups
        self.block_qubits = block_qubits


_PLAN_CACHE: Dict[Any, FusionPlan] = {}


def plan_structure(gate_qubits: Sequence[Tuple[int, ...]], kmax: int) -> FusionPlan:
    """Greedy fusion.  A gate joins the latest block that touches any of its qubits when the
    union stays within ``kmax`` qubits (nothing later touches those qubits, so order is
    preserved); a gate on untouched qubits joins the most recent block that has room."""
    key = (tuple(gate_qubits), kmax)
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        return hit
    groups: List[List[int]] = []
    bq: List[set] = []
    last: Dict[int, int] = {}  # qubit -> index of the latest block touching it
    for gi, qs in enumerate(gate_qubits):
        if len(qs) > MAX_BLOCK_K:
            raise ValueError("gate on %d qubits exceeds the widest supported block (%d)" % (len(qs), MAX_BLOCK_K))
        if len(qs) > kmax:  # wider than the fusion cap: a block of its own
            groups.append([gi])
            bq.append(set(qs))
            for q in qs:
                last[q] = len(groups) - 1
            continue
        deps = [last[q] for q in qs if q in last]
        target = -1
        # the gate must run after block b = latest block touching its qubits; blocks after b
        # do not touch them, so it commutes into any of them: first fit, preferring b itself
        b = max(deps) if deps else max(0, len(groups) - 16)
        sq = set(qs)
        for j in range(b, len(groups)):
            if len(bq[j] | sq) <= kmax:
                target = j
                break
            if j - b > 16:
                break
        if target < 0:
            groups.append([])
            bq.append(set())
            target = len(groups) - 1
        groups[target].append(gi)
        bq[target] |= set(qs)
        for q in qs:
            last[q] = target
    plan = FusionPlan(groups, [tuple(sorted(s)) for s in bq])
    if len(_PLAN_CACHE) > 256:
        _PLAN_CACHE.clear()
    _PLAN_CACHE[key] = plan
    return plan


def fuse(ops: Sequence[GateOp], nqubits: int, kmax: int = 4) -> List[Block]:
    """Fuse ``ops`` (in program order) into blocks for a state of ``nqubits`` qubits."""
    if not ops:
        return []
    plan = plan_structure([op.qubits for op in ops], kmax)
    blocks: List[Block] = []
    for grp, qs in zip(plan.groups, plan.block_qubits):
        k = len(qs)
        D = 1 << k
        m: np.ndarray = np.eye(D, dtype=np.complex128)
        qlist = list(qs)
        for gi in grp:
            op = ops[gi]
            m = embed_apply(m, qlist, _raw(op.matrix), list(op.qubits))
        batched = m.ndim == 3
        bits = tuple(sorted(nqubits - 1 - q for q in qs))
        # row/col index: big-endian over ascending qubits == bit j of the index <-> bits[j]
        diag = bool(np.count_nonzero(m - (np.einsum("...ii->...i", m)[..., None] * np.eye(D))) == 0) if not batched else False
        blocks.append(Block(qubits=qs, bits=bits, matrix=m, batched=batched, ngates=len(grp), diagonal=diag))
    return blocks


# ------------------------------------------------------------------------------------------------
# structure-aware fusion (diagonal / permutation / dense gate classes)
# ------------------------------------------------------------------------------------------------
# The gate pass (csrc/lpass.cu, tcb200_apply_gate_pass) treats three classes of matrices
# differently: permutation matrices of an affine bit map (x, cnot, swap ...; any monomial matrix
# on <= 2 qubits once its phases are split off) cost nothing -- they only rewrite the tile's
# index map; diagonal matrices (z, s, t, rz, cz, rzz, cphase ...) cost one table multiply that
# consecutive diagonals share; everything else is a dense 2^k x 2^k multiply (2 * 2^k packed
# FMA per amplitude).  Fusion therefore merges gates only when the merged block is not dearer
# than its parts: r(a) r(b) cnot(a,b) r'(a) r'(b) becomes ONE 4x4 (8 packed FMA per amplitude
# for four 1-qubit gates' worth of work), a bare cnot stays a free permutation, an rzz next to
# an rx stays a shared table instead of turning the pair into a dense 4x4.
# Classes follow tensorcircuit/gates.py:46-127 (constant matrices), :579-636 (rx/ry/rz),
# :826-865 (rzz/rxx/ryy via exponential_gate_unity).

KIND_DENSE, KIND_DIAG, KIND_PERM, KIND_MONO = "dense", "diag", "perm", "mono"
# planner-only class: a dense 1-qubit gate whose elements are each purely real or purely imaginary (rx, ry,
# h ...): the gate pass applies it with half the FMAs (LOP_RA / LOP_RB).  Blocks carry KIND_DENSE -- the
# library recognises the pattern from the matrix itself.
KIND_HALF = "half"


# entries below this magnitude are rounding residue of the gate formulas (cos(pi/2) = 6e-17 in
# iswap, rx(pi) ...): they are treated as exact zeros -- six orders below the complex128 tolerance
SNAP_EPS = 1e-15


def matrix_kind(m: Any) -> str:
    """Class of a gate matrix by its zero pattern (batched: the union over the batch)."""
    a = _raw(m)
    if a.ndim == 2 and a.shape[0] <= 4:
        return _small_matrix_kind(a.tolist())
    if a.ndim == 3 and a.shape[-1] == 2:
        off = np.abs(a[:, 0, 1]).max() > SNAP_EPS or np.abs(a[:, 1, 0]).max() > SNAP_EPS
        if off and not (a[:, 0, 0].imag.any() or a[:, 1, 1].imag.any()):
            if not (a[:, 0, 1].imag.any() or a[:, 1, 0].imag.any()) or not (a[:, 0, 1].real.any() or a[:, 1, 0].real.any()):
                dg = np.abs(a[:, 0, 0]).max() > SNAP_EPS or np.abs(a[:, 1, 1]).max() > SNAP_EPS
                if dg:
                    return KIND_HALF
    mag = np.abs(a)
    if a.ndim == 3:
        mag = mag.max(axis=0)
    nz = mag > SNAP_EPS
    D = nz.shape[-1]
    if not nz[_OFFDIAG[D]].any():
        return KIND_DIAG
    if D <= 4 and (nz.sum(axis=0) == 1).all() and (nz.sum(axis=1) == 1).all():
        vals = a[..., nz]
        return KIND_PERM if np.all(np.abs(vals - 1) <= SNAP_EPS) else KIND_MONO
    return KIND_DENSE


def _small_matrix_kind(rows: List[List[complex]]) -> str:
    """matrix_kind for a 2x2 / 4x4 matrix given as nested lists (no numpy call overhead)"""
    D = len(rows)
    diag = True
    colcount = [0] * D
    mono = True
    ones = True
    for i in range(D):
        cnt = 0
        ri = rows[i]
        for j in range(D):
            v = ri[j]
            if abs(v) > SNAP_EPS:
                cnt += 1
                colcount[j] += 1
                if i != j:
                    diag = False
                if abs(v - 1) > SNAP_EPS:
                    ones = False
        if cnt != 1:
            mono = False
    if diag:
        return KIND_DIAG
    if mono and all(c == 1 for c in colcount):
        return KIND_PERM if ones else KIND_MONO
    if D == 2:
        (a, b), (c, d) = rows
        if a.imag == 0 and d.imag == 0 and ((b.imag == 0 and c.imag == 0) or (b.real == 0 and c.real == 0)):
            return KIND_HALF
    return KIND_DENSE


_OFFDIAG = {1 << k: ~np.eye(1 << k, dtype=bool) for k in range(0, 6)}


def _kind_cost(kind: str, k: int) -> int:
    """packed FMA per amplitude the gate pass spends on a block of this class"""
    if kind == KIND_DENSE:
        return 2 << k
    if kind == KIND_PERM:
        return 0
    return 2  # one table multiply / a half-cost 1-qubit gate


def _combine_kinds(kinds: Sequence[str]) -> str:
    if KIND_DENSE in kinds or KIND_HALF in kinds:
        return KIND_DENSE  # (a product of half-cost gates may be half-cost again: fuse_structured looks at the matrix)
    if all(k == KIND_DIAG for k in kinds):
        return KIND_DIAG
    if all(k == KIND_PERM for k in kinds):
        return KIND_PERM
    return KIND_MONO


def plan_structure_kinds(gate_qubits: Sequence[Tuple[int, ...]], gate_kinds: Sequence[str], kmax: int,
                          gate_batched: Optional[Sequence[bool]] = None) -> FusionPlan:
    """Cost-aware greedy fusion.  A gate merges with the blocks that are currently the latest on
    its qubits when (i) each of them is still the latest block on ALL of its own qubits (so it
    commutes forward to the merge point), (ii) the union stays within ``kmax`` qubits (2 for
    monomial / diagonal results) and (iii) the merged block costs no more than the parts."""
    # vmap: a merge that involves per-element matrices costs a batched matrix product on the host
    # for every call; it is only taken when it saves device work by a clear margin
    gb = tuple(bool(x) for x in gate_batched) if gate_batched is not None else (False,) * len(gate_qubits)
    key = ("kinds", tuple(gate_qubits), tuple(gate_kinds), kmax, gb if any(gb) else None)
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        return hit
    groups: List[List[int]] = []
    bq: List[set] = []
    bkind: List[str] = []
    bbat: List[bool] = []
    last: Dict[int, int] = {}
    # next gate on each qubit after gate gi (lookahead credit of the inner merge below)
    ng = len(gate_qubits)
    nxt_on: List[Dict[int, int]] = [dict() for _ in range(ng)]
    seen: Dict[int, int] = {}
    for gi in range(ng - 1, -1, -1):
        for q in gate_qubits[gi]:
            if q in seen:
                nxt_on[gi][q] = seen[q]
        for q in gate_qubits[gi]:
            seen[q] = gi

    def merge_into(target: int, others: Sequence[int], union: set, kind_m: str) -> None:
        for b in others:  # earlier blocks commute forward into the merge point
            groups[target] = groups[b] + groups[target]
            groups[b] = []
            bq[b] = set()
        groups[target] = sorted(groups[target])
        bq[target] = union
        bkind[target] = kind_m
        bbat[target] = bbat[target] or any(bbat[b] for b in others)

    for gi, (qs, kg) in enumerate(zip(gate_qubits, gate_kinds)):
        if len(qs) > MAX_BLOCK_K:
            raise ValueError("gate on %d qubits exceeds the widest supported block (%d)" % (len(qs), MAX_BLOCK_K))
        sq = set(qs)
        tops = sorted({last[q] for q in qs if q in last})
        ontop = [b for b in tops if all(last[q] == b for q in bq[b])]
        target = -1
        if tops and len(ontop) == len(tops):
            # full merge: every block that is the latest on one of the gate's qubits joins
            union = set(sq)
            for b in tops:
                union |= bq[b]
            kind_m = _combine_kinds([bkind[b] for b in tops] + [kg])
            limit = kmax if kind_m == KIND_DENSE else 2
            parts = sum(_kind_cost(bkind[b], len(bq[b])) for b in tops) + _kind_cost(kg, len(qs))
            margin = 3 if (gb[gi] or any(bbat[b] for b in tops)) and kind_m == KIND_DENSE and len(union) > 1 else 0
            if len(union) <= max(limit, len(qs)) and len(union) <= MAX_BLOCK_K and _kind_cost(kind_m, len(union)) + margin <= parts:
                target = tops[-1]
                merge_into(target, tops[:-1], union, kind_m)
        if target < 0 and len(qs) == 2 and len(qs) <= kmax:
            # inner merge: only the blocks that live entirely on the gate's own qubits join (the
            # other latest blocks simply stay in front).  Worth it when a dense 1-qubit block is
            # upgraded to a dense 2-qubit block that the following 1-qubit gates then join for
            # free (lookahead: 4 packed FMA of credit per qubit whose next gate is such a gate).
            inner = [b for b in ontop if bq[b] <= sq]
            if inner and any(bkind[b] in (KIND_DENSE, KIND_HALF) for b in inner):
                kind_m = _combine_kinds([bkind[b] for b in inner] + [kg])
                parts = sum(_kind_cost(bkind[b], len(bq[b])) for b in inner) + _kind_cost(kg, len(qs))
                credit = 0
                for q in qs:
                    j = nxt_on[gi].get(q)
                    if j is not None and len(gate_qubits[j]) == 1 and gate_kinds[j] in (KIND_DENSE, KIND_HALF):
                        credit += _kind_cost(gate_kinds[j], 1)
                margin = 3 if (gb[gi] or any(bbat[b] for b in inner)) else 0
                if _kind_cost(kind_m, len(sq)) + margin <= parts + credit:
                    # the merged block holds the gate, which must follow the latest blocks that
                    # are NOT merged: it becomes a new block at the end, the inner ones move into it
                    groups.append([])
                    bq.append(set())
                    bkind.append(kind_m)
                    bbat.append(False)
                    target = len(groups) - 1
                    merge_into(target, inner, set(sq), kind_m)
        if target < 0:
            groups.append([])
            bq.append(set(sq))
            bkind.append(kg)
            bbat.append(False)
            target = len(groups) - 1
        groups[target].append(gi)
        bbat[target] = bbat[target] or gb[gi]
        for q in bq[target]:
            last[q] = target
    keep = [i for i in range(len(groups)) if groups[i]]
    plan = FusionPlan([groups[i] for i in keep], [tuple(sorted(bq[i])) for i in keep])
    if len(_PLAN_CACHE) > 256:
        _PLAN_CACHE.clear()
    _PLAN_CACHE[key] = plan
    return plan


def fuse_structured(ops: Sequence[GateOp], nqubits: int, kmax: int = 2) -> List[Block]:
    """Structure-aware fusion of ``ops`` (program order) for the gate pass: blocks carry ``kind``
    ('dense' / 'diag' / 'perm' / 'mono') computed from the fused matrix."""
    if not ops:
        return []
    kinds = [op.kind if op.kind is not None else matrix_kind(op.matrix) for op in ops]
    plan = plan_structure_kinds([op.qubits for op in ops], kinds, kmax, [is_batched(op.matrix) for op in ops])
    blocks: List[Block] = []
    for grp, qs in zip(plan.groups, plan.block_qubits):
        k = len(qs)
        D = 1 << k
        qlist = list(qs)
        single = len(grp) == 1 and tuple(ops[grp[0]].qubits) == tuple(qs)
        if single:
            m = np.asarray(_raw(ops[grp[0]].matrix), dtype=np.complex128)
        else:
            m = np.eye(D, dtype=np.complex128)
            for gi in grp:
                op = ops[gi]
                m = embed_apply(m, qlist, _raw(op.matrix), list(op.qubits))
        batched = m.ndim == 3
        bits = tuple(sorted(nqubits - 1 - q for q in qs))
        kind = kinds[grp[0]] if single else matrix_kind(m)
        if kind == KIND_HALF:
            kind = KIND_DENSE  # planner-only class; the exact zeros of analytic matrices survive as they are
        if kind != KIND_DENSE and not batched:
            # hand the library exact zeros / ones: it classifies by exact comparison
            m = np.where(np.abs(m) > SNAP_EPS, m, 0)
            if kind == KIND_PERM:
                m = np.where(m != 0, 1.0 + 0j, 0)
        blocks.append(Block(qubits=qs, bits=bits, matrix=m, batched=batched, ngates=len(grp), diagonal=kind == KIND_DIAG, kind=kind))
    return blocks


def clear_plan_cache() -> None:
    _PLAN_CACHE.clear()
    _PASS_CACHE.clear()


# ------------------------------------------------------------------------------------------------
# pass planning: several fused blocks per HBM read + write
# ------------------------------------------------------------------------------------------------
def tile_hi_fixpoint(bits: Sequence[int], tile_bits: int, nbits: int) -> List[int]:
    """Bits of ``bits`` that do not fall into the contiguous low part [0, tile_bits - h) of a
    tile gathering h high bits (the geometry make_geom_hi builds on the device side)."""
    if nbits <= tile_bits:
        return []
    h = 0
    while True:
        c = sum(1 for b in set(bits) if b >= tile_bits - h)
        if c == h:
            break
        h = c
    return sorted(b for b in set(bits) if b >= tile_bits - h)


@dataclass
class Pass:
    """Blocks executed on one staged tile: one HBM read + write of the state."""

    block_ids: List[int]
    tile_hi: List[int]


@dataclass
class RegTile:
    """<= 4 tile-local bits and the blocks (1 or 2 bits each, all inside those bits) applied to
    them back to back in registers: one shared-memory round trip for all of them."""

    bits: Tuple[int, ...]
    block_ids: List[int]


def plan_regtiles(block_bits: Sequence[Tuple[int, ...]], ids: Sequence[int], max_bits: int = 4,
                  max_gates: int = 12) -> List[RegTile]:
    """Cluster the blocks of one pass (``ids`` in execution order) into register tiles.

    A block may join the open register tile when every earlier block of the pass that shares a
    bit with it is already placed (dependencies), and the union of bits stays within
    ``max_bits``; blocks on the same qubits therefore pile up in one tile."""
    n = len(ids)
    bsets = [set(block_bits[i]) for i in ids]
    placed = [False] * n
    tiles: List[RegTile] = []
    left = n
    while left:
        bits: set = set()
        members: List[int] = []
        progressed = True
        while progressed and len(members) < max_gates:
            progressed = False
            for i in range(n):
                if placed[i]:
                    continue
                if any((not placed[j]) and (bsets[j] & bsets[i]) for j in range(i)):
                    continue  # an earlier block on a shared bit is still pending
                if len(bits | bsets[i]) > max_bits:
                    continue
                bits |= bsets[i]
                members.append(i)
                placed[i] = True
                left -= 1
                progressed = True
                break
        tiles.append(RegTile(bits=tuple(sorted(bits)), block_ids=[ids[i] for i in members]))
    return tiles


_PASS_CACHE: Dict[Any, List[Pass]] = {}


def plan_passes(block_bits: Sequence[Tuple[int, ...]], nbits: int, tile_bits: int, max_hi: int = 6,
                max_ops: int = 16, max_mat_elems: int = 1536, max_pass_k: int = 4, nseeds: int = 8,
                block_cost: Optional[Sequence[int]] = None, block_weight: Optional[Sequence[float]] = None,
                jitter_seed: Optional[int] = None, block_diag: Optional[Sequence[bool]] = None) -> List[Pass]:
    """List scheduling of fused blocks into tile passes, with a one-pass lookahead.

    A tile holds the ``tile_bits - h`` lowest index bits plus ``h <= max_hi`` gathered high bits;
    a block can run in a pass when all its bits are in the tile.  Blocks are taken in dependency
    order (a block is ready when every earlier block sharing a bit with it is done -- possibly
    earlier in the same pass).  A pass is filled greedily -- among ready blocks the one that
    needs the fewest new gathered bits goes first -- but the greedy fill is tried from each of
    the first ``nseeds`` ready blocks as its opening block, and the fullest pass wins (ties: the
    earliest seed).  On the config-4 recipe at n = 34 that gives 37 passes instead of 45.  The
    plan depends only on the bit structure and is cached."""
    max_hi = max(0, min(max_hi, tile_bits - 4))  # keep rows of >= 16 amplitudes contiguous
    low_fixed = tile_bits - max_hi              # bits below this are in every tile
    # block_cost: parameter-bank elements per block (default 4^k); block_weight: what "fullest
    # pass" counts (default 1 per block; the gate pass gives free permutation blocks weight 0)
    cost = list(block_cost) if block_cost is not None else [1 << (2 * len(b)) for b in block_bits]
    weight = list(block_weight) if block_weight is not None else [1.0] * len(block_bits)
    # jitter_seed: break the greedy's ties (equal number of new gathered bits) at random instead of
    # in program order -- plan_passes_best keeps the best of several such plans
    jit = np.random.default_rng(jitter_seed) if jitter_seed is not None else None
    # block_diag: diagonal blocks commute with each other, so a diagonal block is only ordered against the
    # non-diagonal blocks around it -- an rzz ladder is 27 independent blocks, not a chain from qubit 0 to n-1
    # (any order consistent with this relaxed DAG gives the same product up to rounding)
    diag = [bool(x) for x in block_diag] if block_diag is not None else [False] * len(block_bits)
    key = (tuple(block_bits), nbits, tile_bits, max_hi, max_ops, max_mat_elems, max_pass_k, nseeds, tuple(cost), tuple(weight), jitter_seed,
           tuple(diag) if any(diag) else None)
    hit = _PASS_CACHE.get(key)
    if hit is not None:
        return hit
    nb = len(block_bits)
    preds: List[set] = [set() for _ in range(nb)]
    succs: List[List[int]] = [[] for _ in range(nb)]
    last: Dict[int, int] = {}           # per bit: the last non-diagonal block
    dsince: Dict[int, List[int]] = {}   # per bit: the diagonal blocks after it
    for i, bits in enumerate(block_bits):
        for q in bits:
            if diag[i]:
                if q in last:
                    preds[i].add(last[q])
                dsince.setdefault(q, []).append(i)
            else:
                if dsince.get(q):
                    preds[i].update(dsince[q])
                    dsince[q] = []
                elif q in last:
                    preds[i].add(last[q])
                last[q] = i
    for i in range(nb):
        for p in preds[i]:
            succs[p].append(i)
    indeg0 = [len(p) for p in preds]
    ready0 = [i for i in range(nb) if indeg0[i] == 0]

    def fill(seed: int, indeg_in: List[int], ready_in: List[int]) -> Tuple[List[int], set, List[int], List[int]]:
        """one pass opened by ``seed``: (blocks, bits used, indeg and ready list afterwards)"""
        indeg, ready = list(indeg_in), list(ready_in)
        cur: List[int] = []
        used: set = set()
        cur_hi: List[int] = []
        mat = 0
        while len(cur) < max_ops:
            best, best_score, best_hi = -1, None, None
            for i in ([seed] if not cur else ready):
                bits = block_bits[i]
                k = len(bits)
                if k > max_pass_k:
                    if not cur:  # too wide for the staged pass: runs alone through the dense kernel
                        best, best_hi, best_score = i, [], (0, i)
                        break
                    continue
                if mat + cost[i] > max_mat_elems:
                    continue
                hi = tile_hi_fixpoint(list(used | set(bits)), tile_bits, nbits)
                if len(hi) > max_hi:
                    continue
                # cost of a candidate = tile bits it adds above the always-present low bits.  (Counting the growth of
                # the gathered set instead makes bits just above the low part look free while few bits are gathered,
                # and they turn into gathered bits later: 35 instead of 32 passes on the config-4 recipe at n = 34.)
                nnew = sum(1 for b in set(bits) if b >= low_fixed and b not in used)
                score = (nnew, i if jit is None else float(jit.random()))
                if best_score is None or score < best_score:
                    best, best_score, best_hi = i, score, hi
                    if score[0] <= 0 and jit is None:
                        break
            if best < 0:
                if cur:
                    break
                # the seed fits no tile (more high bits than one can gather): it runs alone through
                # the single-block kernel, which chooses its own geometry
                best, best_hi = seed, []
                standalone = True
            else:
                standalone = len(block_bits[best]) > max_pass_k
            cur.append(best)
            used |= set(block_bits[best])
            cur_hi = best_hi
            mat += cost[best]
            ready.remove(best)
            for s_ in succs[best]:
                indeg[s_] -= 1
                if indeg[s_] == 0:
                    ready.append(s_)
            ready.sort()
            if standalone:
                break
        return cur, used, indeg, ready

    indeg, ready = indeg0, sorted(ready0)
    passes: List[Pass] = []
    remaining = nb
    while remaining:
        best_fill = None
        for seed in ready[: max(1, nseeds)]:
            f = fill(seed, indeg, ready)
            if best_fill is None or sum(weight[i] for i in f[0]) > sum(weight[i] for i in best_fill[0]):
                best_fill = f
        if best_fill is None or not best_fill[0]:
            raise RuntimeError("pass planner made no progress")
        cur, used, indeg, ready = best_fill
        remaining -= len(cur)
        passes.append(Pass(block_ids=cur, tile_hi=tile_hi_fixpoint(list(used), tile_bits, nbits)))
    if len(_PASS_CACHE) > 64:
        _PASS_CACHE.clear()
    _PASS_CACHE[key] = passes
    return passes
