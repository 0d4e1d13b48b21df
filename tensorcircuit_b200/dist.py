"""State vector sharded over G = 2^g GPUs (north-star item 6).

Layout: rank r holds the contiguous slice of amplitudes whose top g *physical* index bits equal
r -- with the identity qubit map these are TC qubits 0..g-1, i.e. exactly the reference's flat
big-endian vector split G ways.  Local kernels run unchanged on the n-g local bits.

A fused block whose matrix is block-diagonal on its global bits (controls, rz/rzz/cz/phase on a
global qubit) needs no data movement: the rank picks the sub-matrix selected by its own bits.
Any other block that touches a global qubit triggers a *remap*: the g global bits are swapped
with the top g local bits by one NCCL all-to-all -- every peer chunk is then contiguous, so no
pack kernel is needed, and a logical->physical bit table is updated instead of moving data
back.  The scheduler executes every ready block it can between remaps (blocks that commute
are reordered across the remap boundary) to amortise the NVLink traffic: (1 - 2^-g) of the
shard per remap against one HBM pass per block.

Two exchange modes: double-buffered (one ``all_to_all_single`` into a second shard-sized
buffer, pointers swapped) when memory allows, otherwise chunked through a bounded staging
buffer (for 128 GiB shards on a 180 GB device).

One process per GPU (torchrun); ``torch.distributed`` is the plumbing (NCCL on GPUs; the CPU
tests run the same scheduler over gloo with an emulated local engine)."""

from __future__ import annotations

import math
import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .fusion import Block, matrix_kind

_SWAP = np.eye(4, dtype=np.complex128)[[0, 2, 1, 3]]


def permute_matrix_bits(m: np.ndarray, order: Sequence[int]) -> np.ndarray:
    """New index bit j' <- old index bit order[j'] (row and column), batch axis allowed."""
    k = len(order)
    lead = m.ndim - 2
    t = m.reshape(m.shape[:lead] + (2,) * (2 * k))
    # axis a (big-endian) <-> index bit k-1-a
    src_axes = [k - 1 - order[k - 1 - a] for a in range(k)]
    perm = list(range(lead)) + [lead + a for a in src_axes] + [lead + k + a for a in src_axes]
    return np.ascontiguousarray(np.transpose(t, perm)).reshape(m.shape)


def restrict_global(m: np.ndarray, is_global: Sequence[bool], gvals: Sequence[int], tol: float = 0.0) -> Optional[np.ndarray]:
    """If ``m`` (index bit j flagged by is_global[j]) is block diagonal w.r.t. its global bits,
    return the block selected by the global bit values ``gvals`` (one per flagged bit, in bit
    order); otherwise None."""
    k = len(is_global)
    D = 1 << k
    gmask = sum(1 << j for j in range(k) if is_global[j])
    idx = np.arange(D)
    off = (idx[:, None] & gmask) != (idx[None, :] & gmask)
    if np.any(np.abs(m[..., off]) > tol):
        return None
    sel = 0
    gi = 0
    for j in range(k):
        if is_global[j]:
            sel |= (int(gvals[gi]) & 1) << j
            gi += 1
    keep = idx[(idx & gmask) == sel]
    return np.ascontiguousarray(m[..., keep[:, None], keep[None, :]])


class DistState:
    """A 2^n_total state sharded over the ranks of ``group``."""

    def __init__(self, n_total: int, dtype: str = "complex64", group: Any = None,
                 state_factory: Optional[Callable[..., Any]] = None, staging_bytes: int = 8 << 30,
                 double_buffer: Optional[bool] = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.G = dist.get_world_size(group)
        self.g = int(round(math.log2(self.G)))
        if 1 << self.g != self.G:
            raise ValueError("world size must be a power of two")
        self.n = int(n_total)
        self.nloc = self.n - self.g
        if self.nloc < 2 * self.g:
            raise ValueError("state too small to shard over %d ranks" % self.G)
        self.dtype = dtype
        if state_factory is None:
            from .engine import DeviceState

            state_factory = DeviceState
        self._factory = state_factory
        self.local = state_factory(self.nloc, dtype)
        self.amp_bytes = 8 if dtype == "complex64" else 16
        self.shard_bytes = self.amp_bytes << self.nloc
        self.is_cuda = self.local.buf.is_cuda
        if double_buffer is None:
            if self.is_cuda:
                free, _ = torch.cuda.mem_get_info()
                double_buffer = free > self.shard_bytes + (4 << 30)
            else:
                double_buffer = True
        self.alt = state_factory(self.nloc, dtype) if (double_buffer and self.G > 1) else None
        self.staging_bytes = int(staging_bytes)
        self._stage: Optional[torch.Tensor] = None
        # logical bit -> physical bit (physical bits >= nloc are the rank bits)
        self.phys: List[int] = list(range(self.n))
        self.stats = {"remaps": 0, "remap_bytes": 0, "local_passes": 0, "swap_passes": 0, "remap_ms": 0.0}
        self._remap_events: List[Any] = []

    # -- bookkeeping -------------------------------------------------------------------------
    def logical_at(self, p: int) -> int:
        return self.phys.index(p)

    def is_global_bit(self, b: int) -> bool:
        return self.phys[b] >= self.nloc

    def rank_bit(self, p: int) -> int:
        return (self.rank >> (p - self.nloc)) & 1

    # -- initial state -------------------------------------------------------------------------
    def init_zero(self) -> None:
        self.phys = list(range(self.n))
        if self.rank == 0:
            self.local.init_zero()
        else:
            self._zero(self.local)

    def _zero(self, st: Any) -> None:
        if self.is_cuda:
            from . import engine
            from ._lib import check, lib

            check(lib.tcb200_set_zero(engine._ptr(st.buf), st.nbits, st.dt, 1, engine._stream()))
        else:
            st.buf.zero_()

    # -- local execution ---------------------------------------------------------------------------
    def _try_local(self, blk: Block) -> Optional[Block]:
        """The block rewritten on physical local bits, or None if it needs a remap."""
        pb = [self.phys[b] for b in blk.bits]
        isg = [p >= self.nloc for p in pb]
        m = blk.matrix
        if any(isg):
            gv = [self.rank_bit(p) for p in pb if p >= self.nloc]
            m = restrict_global(m, isg, gv)
            if m is None:
                return None
            pb = [p for p in pb if p < self.nloc]
            if not pb:  # pure phase on global bits: scalar
                return Block(qubits=(), bits=(), matrix=m, batched=False, ngates=blk.ngates)
        order = list(np.argsort(pb))
        m = permute_matrix_bits(m, order)
        bits = tuple(sorted(pb))
        return Block(qubits=tuple(self.nloc - 1 - b for b in bits), bits=bits, matrix=m, batched=False, ngates=blk.ngates,
                     kind=matrix_kind(m))

    def _run_local(self, lb: Block) -> None:
        if len(lb.bits) == 0:
            # global phase for this rank: fold into a 1-bit diagonal block
            ph = complex(np.asarray(lb.matrix).reshape(-1)[0])
            lb = Block(qubits=(self.nloc - 1,), bits=(0,), matrix=np.eye(2, dtype=np.complex128) * ph, batched=False, ngates=lb.ngates, kind="diag")
        self.local.apply_block(lb)
        self.stats["local_passes"] += 1

    def _run_local_batch(self, batch: Sequence[Block]) -> None:
        """Blocks that can all run before the next remap: staged multi-block passes on the
        shard when the local engine plans them, one pass per block otherwise."""
        if not batch:
            return
        fixed = []
        for lb in batch:
            if len(lb.bits) == 0:  # pure rank phase: fold into a 1-bit diagonal block
                ph = complex(np.asarray(lb.matrix).reshape(-1)[0])
                lb = Block(qubits=(self.nloc - 1,), bits=(0,), matrix=np.eye(2, dtype=np.complex128) * ph, batched=False, ngates=lb.ngates, kind="diag")
            fixed.append(lb)
        if self.use_passes and hasattr(self.local, "apply_planned"):
            self.stats["local_passes"] += int(self.local.apply_planned(fixed))
        else:
            for lb in fixed:
                self.local.apply_block(lb)
            self.stats["local_passes"] += len(fixed)

    use_passes = True

    def _swap_local_bits(self, pa: int, pb: int) -> None:
        """Exchange two physical local bits (one local pass) and update the map."""
        lo, hi = sorted((pa, pb))
        self.local.apply_block(Block(qubits=(self.nloc - 1 - hi, self.nloc - 1 - lo), bits=(lo, hi), matrix=_SWAP, batched=False, ngates=0, kind="perm"))
        la, lb = self.logical_at(pa), self.logical_at(pb)
        self.phys[la], self.phys[lb] = pb, pa
        self.stats["swap_passes"] += 1

    def finalize_stats(self) -> Dict[str, Any]:
        """Fold the CUDA-event times of the remaps issued so far into stats['remap_ms'] (synchronises
        the device once; the data path itself never waits for the host)."""
        if self._remap_events:
            torch.cuda.synchronize()
            for t0, t1 in self._remap_events:
                self.stats["remap_ms"] += t0.elapsed_time(t1)
            self._remap_events = []
        return self.stats

    # -- the exchange ---------------------------------------------------------------------------------
    def remap(self) -> None:
        """Swap the g global bits with the top g local bits (all-to-all over contiguous chunks)."""
        if self.G == 1:
            return
        t0 = None
        if self.is_cuda:
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
        src = torch.view_as_real(self.local.buf[0])
        if self.alt is not None:
            dst = torch.view_as_real(self.alt.buf[0])
            dist.all_to_all_single(dst, src, group=self.group)
            self.local, self.alt = self.alt, self.local
        else:
            self._remap_chunked(src)
        if t0 is not None:
            t1.record()
            self._remap_events.append((t0, t1))  # read in finalize_stats(): no host sync on the data path
        for j in range(self.g):
            a, b = self.logical_at(self.nloc - self.g + j), self.logical_at(self.nloc + j)
            self.phys[a], self.phys[b] = self.nloc + j, self.nloc - self.g + j
        self.stats["remaps"] += 1
        self.stats["remap_bytes"] += self.shard_bytes - self.shard_bytes // self.G

    def _remap_chunked(self, src: torch.Tensor) -> None:
        """In-place exchange through a bounded staging buffer (shards too large to double
        buffer).  The rank's own chunk never moves; the peer chunks go slice by slice through two
        staging sets so that packing slice i+1 and unpacking slice i-1 (device copies, HBM-bound)
        overlap the all-to-all of slice i (NVLink-bound)."""
        G, r = self.G, self.rank
        chunk = src.shape[0] // G  # amplitudes per peer chunk
        nset = 2
        per = max(1, min(chunk, self.staging_bytes // (2 * nset * (G - 1) * self.amp_bytes)))
        per = 1 << (per.bit_length() - 1)  # chunk is a power of two
        rows = (G - 1) * per
        if self._stage is None or tuple(self._stage.shape[:3]) != (nset, 2, rows):
            self._stage = torch.empty((nset, 2, rows, 2), dtype=src.dtype, device=src.device)
        peers = [c for c in range(G) if c != r]
        splits = [per] * G
        splits[r] = 0
        slices = list(range(0, chunk, per))
        works: List[Any] = [None] * nset

        def unpack(i: int) -> None:
            k = i % nset
            works[k].wait()
            recv = self._stage[k, 1]
            for j, c in enumerate(peers):
                self._copy_rows(src[c * chunk + slices[i]:], per, recv[j * per:], per, per, 1)

        for i, s in enumerate(slices):
            k = i % nset
            if i >= nset:
                unpack(i - nset)  # frees staging set k
            send = self._stage[k, 0]
            for j, c in enumerate(peers):
                self._copy_rows(send[j * per:], per, src[c * chunk + s:], per, per, 1)
            works[k] = dist.all_to_all_single(self._stage[k, 1], send, output_split_sizes=splits, input_split_sizes=splits,
                                              group=self.group, async_op=True)
        for i in range(max(0, len(slices) - nset), len(slices)):
            unpack(i)

    def _copy_rows(self, dst: torch.Tensor, dst_pitch: int, src: torch.Tensor, src_pitch: int, row: int, nrows: int) -> None:
        """copy nrows runs of `row` amplitudes; pitches in amplitudes"""
        if self.is_cuda:
            from . import engine
            from ._lib import check, lib

            b = self.amp_bytes
            check(lib.tcb200_copy_rows(engine._ptr(dst), dst_pitch * b, engine._ptr(src), src_pitch * b, row * b, nrows, engine._stream()))
        else:
            for r in range(nrows):
                dst[r * dst_pitch : r * dst_pitch + row].copy_(src[r * src_pitch : r * src_pitch + row])

    # -- scheduler --------------------------------------------------------------------------------------
    def run(self, blocks: Sequence[Block]) -> None:
        """Execute fused blocks (logical bits, program order) with as few remaps as the
        dependency structure allows."""
        nb = len(blocks)
        preds: List[set] = [set() for _ in range(nb)]
        succs: List[List[int]] = [[] for _ in range(nb)]
        last: Dict[int, int] = {}
        for i, b in enumerate(blocks):
            for q in b.bits:
                if q in last:
                    preds[i].add(last[q])
                last[q] = i
        for i in range(nb):
            for p in preds[i]:
                succs[p].append(i)
        indeg = [len(p) for p in preds]
        ready = sorted(i for i in range(nb) if indeg[i] == 0)
        done = 0
        done_flag = [False] * nb
        stuck_rounds = 0
        while done < nb:
            progressed = False
            again = True
            batch: List[Block] = []  # everything executable before the next remap, in order
            while again:
                again = False
                for i in list(ready):
                    lb = self._try_local(blocks[i])
                    if lb is None:
                        continue
                    batch.append(lb)
                    ready.remove(i)
                    done += 1
                    done_flag[i] = True
                    progressed = again = True
                    for s in succs[i]:
                        indeg[s] -= 1
                        if indeg[s] == 0:
                            ready.append(s)
                    ready.sort()
            if done < nb and self.G > 1:
                # choose the qubits that become global and move them into the swap window with
                # SWAP blocks that ride at the end of this batch (permutations cost nothing inside
                # a gate pass) instead of one extra pass per moved bit
                batch += self._plan_evictions(blocks, done_flag)
            self._run_local_batch(batch)
            if done == nb:
                break
            if not progressed:
                stuck_rounds += 1
                if stuck_rounds > 4:
                    raise RuntimeError("distributed scheduler made no progress")
            else:
                stuck_rounds = 0
            self.remap()

    @staticmethod
    def _nonlocal_bits(blk: Block) -> Tuple[int, ...]:
        """bits of the block on which its matrix is NOT block diagonal: the ones that must be local"""
        cached = getattr(blk, "_nl", None)
        if cached is not None:
            return cached
        k = len(blk.bits)
        m = np.asarray(blk.matrix)
        out = []
        for j in range(k):
            if restrict_global(m, [x == j for x in range(k)], [0]) is None:
                out.append(blk.bits[j])
        blk._nl = tuple(out)  # type: ignore[attr-defined]
        return blk._nl  # type: ignore[attr-defined]

    def _plan_evictions(self, blocks: Sequence[Block], done_flag: Sequence[bool]) -> List[Block]:
        """Before a remap: the g local qubits whose next non-diagonal use lies farthest ahead go into
        the swap window (they become global).  Returns the SWAP blocks (physical bits) that put them
        there and updates the logical -> physical map accordingly."""
        top0 = self.nloc - self.g
        INF = 1 << 60
        nxt = {q: INF for q in range(self.n)}
        for i, b in enumerate(blocks):
            if done_flag[i]:
                continue
            for q in self._nonlocal_bits(b):
                if nxt[q] == INF:
                    nxt[q] = i
        local_q = [q for q in range(self.n) if self.phys[q] < self.nloc]
        # farthest next use first; ties: qubits already in the window (no swap needed)
        local_q.sort(key=lambda q: (-nxt[q], 0 if self.phys[q] >= top0 else 1))
        evict = local_q[: self.g]
        window_free = [p for p in range(top0, self.nloc) if self.logical_at(p) not in evict]
        swaps: List[Block] = []
        for q in evict:
            pq = self.phys[q]
            if pq >= top0:
                continue
            pw = window_free.pop(0)
            lo, hi = sorted((pq, pw))
            swaps.append(Block(qubits=(self.nloc - 1 - hi, self.nloc - 1 - lo), bits=(lo, hi), matrix=_SWAP, batched=False, ngates=0, kind="perm"))
            other = self.logical_at(pw)
            self.phys[q], self.phys[other] = pw, pq
            self.stats["swap_blocks"] = self.stats.get("swap_blocks", 0) + 1
        return swaps

    def _prepare_remap(self, blocked: Sequence[Block]) -> None:
        """A blocked block that also uses a top-local bit would stay blocked after the swap
        (that bit becomes global): move such bits below the swap window first."""
        top0 = self.nloc - self.g
        protect = set()
        for b in blocked:
            for q in b.bits:
                protect.add(q)
        moving = [q for q in protect if top0 <= self.phys[q] < self.nloc]
        if not moving:
            return
        # candidate low positions: local, below the window, logical bit not needed by blocked blocks
        free = [p for p in range(top0 - 1, -1, -1) if self.logical_at(p) not in protect]
        for q in moving:
            if not free:
                raise RuntimeError("no free local bit to park a qubit before the remap")
            self._swap_local_bits(self.phys[q], free.pop(0))

    def make_local(self, bits: Sequence[int]) -> None:
        """Ensure the logical ``bits`` are local (used by expectation / sampling of X,Y terms)."""
        need = [b for b in bits if self.is_global_bit(b)]
        if not need:
            return
        top0 = self.nloc - self.g
        protect = set(bits)
        moving = [q for q in protect if top0 <= self.phys[q] < self.nloc]
        free = [p for p in range(top0 - 1, -1, -1) if self.logical_at(p) not in protect]
        for q in moving:
            if not free:
                raise RuntimeError("cannot make all requested bits local at once")
            self._swap_local_bits(self.phys[q], free.pop(0))
        self.remap()

    def _swap_local_many(self, pairs: Sequence[Tuple[int, int]]) -> None:
        """A sequence of physical local bit exchanges as ONE planned run of passes: the SWAPs are permutation blocks,
        which the gate pass executes as index remaps, so a whole bit permutation costs a few passes over the shard
        (one per set of bits that fits a tile), not one per exchange."""
        blocks = []
        for pa, pb in pairs:
            if pa == pb:
                continue
            lo, hi = sorted((pa, pb))
            blocks.append(Block(qubits=(self.nloc - 1 - hi, self.nloc - 1 - lo), bits=(lo, hi), matrix=_SWAP, batched=False, ngates=0, kind="perm"))
            la, lb = self.logical_at(pa), self.logical_at(pb)
            self.phys[la], self.phys[lb] = pb, pa
        if not blocks:
            return
        if self.use_passes and hasattr(self.local, "apply_planned"):
            self.stats["swap_passes"] += int(self.local.apply_planned(blocks))
        else:
            for b in blocks:
                self.local.apply_block(b)
            self.stats["swap_passes"] += len(blocks)

    def restore_identity(self) -> None:
        """Bring the state back to the layout it started in (logical bit b at physical bit b: the reference's flat
        vector split G ways): at most two remaps plus the passes of two local bit permutations.  Queries that depend on
        the ORDER of the amplitudes (the CDF of ``sample(status=...)``) are then the single-GPU ones."""
        if self.phys == list(range(self.n)):
            return
        want_top = list(range(self.nloc, self.n))  # logical bits that belong to the rank index, slot j = bit nloc + j
        if self.G > 1 and any(self.phys[b] != b for b in want_top):
            top0 = self.nloc - self.g
            if any(self.is_global_bit(b) for b in want_top):
                # some wanted bits are global in a wrong slot or next to strangers: every global bit comes down first;
                # a wanted bit inside the window would go up in exchange, so it is parked below the window before
                free = [p for p in range(top0 - 1, -1, -1) if self.logical_at(p) not in want_top]
                pairs = []
                for b in want_top:
                    if top0 <= self.phys[b] < self.nloc:
                        if not free:
                            raise RuntimeError("no free local bit to park a qubit before the remap")
                        pairs.append((self.phys[b], free.pop(0)))
                self._swap_local_many(pairs)
                self.remap()
            # all wanted bits are local now: logical bit nloc + j goes to window position top0 + j, then up
            pairs = []
            cur = list(self.phys)
            for j, b in enumerate(want_top):
                src, dst = cur[b], top0 + j
                if src != dst:
                    pairs.append((src, dst))
                    other = cur.index(dst)
                    cur[b], cur[other] = dst, src
            self._swap_local_many(pairs)
            self.remap()
        # local bits: selection by exchanges, all in one planned run
        pairs = []
        cur = list(self.phys)
        for p in range(self.nloc):
            if cur[p] != p:
                src = cur[p]
                pairs.append((src, p))
                other = cur.index(p)
                cur[p], cur[other] = p, src
        self._swap_local_many(pairs)
        assert self.phys == list(range(self.n)), self.phys

    # how sample() orders the CDF: "physical" (no data movement; reproducible for a fixed GPU count) or "logical"
    # (restore_identity() first: the same indices as a single-GPU run for the same uniforms, up to CDF ties)
    sample_order = os.environ.get("TCB200_DIST_SAMPLE_ORDER", "physical")

    # -- queries ----------------------------------------------------------------------------------------
    def norm2(self) -> float:
        v = torch.tensor([float(self.local.norm2()[0])], dtype=torch.float64, device=self.local.buf.device)
        dist.all_reduce(v, group=self.group)
        return float(v.item())

    def masked_norm2(self, mask: int, value: int) -> float:
        """sum of |psi_e|^2 over (e & mask) == value for logical-bit masks: bits that are global at
        the moment select the ranks that contribute, the others become a local mask"""
        lmask = lval = 0
        mine = True
        for q in range(self.n):
            if (int(mask) >> q) & 1:
                p, bit = self.phys[q], (int(value) >> q) & 1
                if p < self.nloc:
                    lmask |= 1 << p
                    lval |= bit << p
                elif ((self.rank >> (p - self.nloc)) & 1) != bit:
                    mine = False
        m = self.local.masked_norm2(lmask, lval) if mine else 0.0
        v = torch.tensor([float(m)], dtype=torch.float64, device=self.local.buf.device)
        dist.all_reduce(v, group=self.group)
        return float(v.item())

    def expectation_terms(self, flips: Sequence[int], signs: Sequence[int], nys: Sequence[int]) -> np.ndarray:
        """<P_t> for logical-bit masks; complex [nterms].  Z-type factors on global bits become
        a per-rank sign; terms that flip a global bit are evaluated after a remap that makes
        their flip bits local."""
        nt = len(flips)
        out = np.zeros(nt, dtype=np.complex128)
        pending = list(range(nt))
        rounds = 0
        while pending:
            now = [t for t in pending if not any(self.is_global_bit(b) for b in range(self.n) if (int(flips[t]) >> b) & 1)]
            if not now:
                rounds += 1
                if rounds > 2 * self.n:
                    raise RuntimeError("cannot localise the flip bits of the remaining Pauli terms")
                t = pending[0]
                self.make_local([b for b in range(self.n) if (int(flips[t]) >> b) & 1])
                continue
            lf, ls, sgn = [], [], []
            for t in now:
                f = s = 0
                neg = 0
                for b in range(self.n):
                    p = self.phys[b]
                    if (int(flips[t]) >> b) & 1:
                        f |= 1 << p
                    if (int(signs[t]) >> b) & 1:
                        if p >= self.nloc:
                            neg ^= self.rank_bit(p)
                        else:
                            s |= 1 << p
                lf.append(f)
                ls.append(s)
                sgn.append(-1.0 if neg else 1.0)
            vals = self.local.expectation_terms(lf, ls, [nys[t] for t in now])[0] * np.asarray(sgn)
            buf = torch.from_numpy(np.stack([vals.real, vals.imag], axis=-1).copy()).to(self.local.buf.device)
            dist.all_reduce(buf, group=self.group)
            r = buf.cpu().numpy()
            out[now] = r[:, 0] + 1j * r[:, 1]
            pending = [t for t in pending if t not in set(now)]
        return out

    def sample(self, uniforms: Any, logical_order: Optional[bool] = None) -> np.ndarray:
        """CDF sampling over the sharded state; returns *logical* basis-state indices (int64).
        Every rank receives all uniforms; each resolves those that fall into its CDF interval.

        By default the CDF runs over the amplitudes in their current PHYSICAL order (rank-major, then the local
        physical bits): after a remap that is a permutation of the logical order, so the same ``status``
        draws from the same distribution but not the same bitstrings as a single-GPU run -- samples are
        reproducible for a fixed GPU count and circuit, not across GPU counts.  ``logical_order=True``
        (or ``DistState.sample_order = "logical"`` / ``tc.set_distributed(True, sample_order="logical")``) restores
        the identity layout first (``restore_identity``: up to two more remaps and a few local passes): the
        CDF is then the reference's (circuit.py:915-935 over the flat vector) and the indices are those of a
        single-GPU run for the same uniforms, up to ties between adjacent CDF values."""
        if logical_order is None:
            logical_order = self.sample_order == "logical"
        if logical_order:
            self.restore_identity()
        dev = self.local.buf.device
        mine = float(self.local.norm2()[0])
        tot = torch.zeros(self.G, dtype=torch.float64, device=dev)
        tot[self.rank] = mine
        dist.all_reduce(tot, group=self.group)
        t = tot.cpu().numpy()
        offset = float(np.sum(t[: self.rank]))
        total = float(np.sum(t))
        loc = self.local.sample(uniforms, cdf_offset=offset, cdf_total=total)
        phys_idx = np.where(loc >= 0, loc + (self.rank << self.nloc), -1).astype(np.int64)
        buf = torch.from_numpy(phys_idx).to(dev)
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=self.group)
        phys_idx = buf.cpu().numpy()
        # rounding at a shard boundary can leave a shot unowned: it goes to the last amplitude of the shard
        # whose CDF interval it falls next to (not to the global last index, which would bias |1..1>)
        lost = np.nonzero(phys_idx < 0)[0]
        if lost.size:
            u = uniforms.reshape(-1)[torch.from_numpy(lost)].cpu().numpy() if isinstance(uniforms, torch.Tensor) else np.asarray(uniforms, dtype=np.float64).reshape(-1)[lost]
            k = np.clip(np.searchsorted(np.cumsum(t), total * (1.0 - np.asarray(u, dtype=np.float64)), side="left"), 0, self.G - 1)
            phys_idx[lost] = (k.astype(np.int64) << self.nloc) | ((1 << self.nloc) - 1)
        # physical -> logical bit order
        same = 0
        for b in range(self.n):
            if self.phys[b] == b:
                same |= 1 << b
        out = phys_idx & np.int64(same)  # bits that sit where they belong move in one operation
        for b in range(self.n):
            if self.phys[b] != b:
                out |= ((phys_idx >> self.phys[b]) & 1) << b
        return out

    def gather_state(self) -> np.ndarray:
        """Full logical state on every rank (tests / small n only)."""
        loc = torch.view_as_real(self.local.buf[0]).contiguous()
        parts = [torch.empty_like(loc) for _ in range(self.G)]
        dist.all_gather(parts, loc, group=self.group)
        full = torch.cat(parts, dim=0).cpu().numpy()
        phys_state = full[:, 0] + 1j * full[:, 1]
        idx = np.arange(1 << self.n, dtype=np.int64)
        pidx = np.zeros_like(idx)
        for b in range(self.n):
            pidx |= ((idx >> b) & 1) << self.phys[b]
        return phys_state[pidx]


class DistEngineState:
    """``DeviceState``-shaped facade over :class:`DistState` so that ``tc.Circuit`` can run SPMD
    under torchrun (``tc.set_distributed(True)``): every rank records the same circuit, the
    state is sharded, queries return identical host values on every rank."""

    def __init__(self, nbits: int, dtype: str = "complex64", batch: int = 1, **kw: Any):
        if batch != 1:
            raise NotImplementedError("a distributed state cannot also be vmapped; shard the batch instead")
        self.ds = DistState(nbits, dtype, **kw)
        self.nbits = nbits
        self.dtype = dtype
        self.batch = 1

    def init_zero(self) -> None:
        self.ds.init_zero()

    def load(self, src: Any) -> None:
        v = np.asarray(src).reshape(-1)
        lo = self.ds.rank << self.ds.nloc
        self.ds.phys = list(range(self.ds.n))
        self.ds.local.load(v[lo : lo + (1 << self.ds.nloc)])

    def apply_blocks(self, blocks: Sequence[Block]) -> None:
        self.ds.run(blocks)

    def norm2(self) -> np.ndarray:
        return np.asarray([self.ds.norm2()])

    def masked_norm2(self, mask: int, value: int) -> float:
        return self.ds.masked_norm2(mask, value)

    def expectation_terms(self, flips: Sequence[int], signs: Sequence[int], nys: Sequence[int]) -> np.ndarray:
        return self.ds.expectation_terms(flips, signs, nys)[None, :]

    def sample(self, uniforms: Any, **kw: Any) -> np.ndarray:
        return self.ds.sample(uniforms)

    @property
    def buf(self) -> torch.Tensor:
        if self.nbits > 28:
            raise NotImplementedError("wavefunction() of a distributed state above 2^28 amplitudes: query it with expectation_ps / sample")
        full = self.ds.gather_state()
        return torch.from_numpy(full.astype(self.dtype))[None, :]

    def probability(self) -> torch.Tensor:
        # the gathered state is a HOST array (n <= 28, debugging / small-state convenience only)
        full = self.ds.gather_state()
        p = (full.real.astype(np.float64) ** 2 + full.imag.astype(np.float64) ** 2)
        return torch.from_numpy(p.astype(np.float32 if self.dtype == "complex64" else np.float64))[None, :]
