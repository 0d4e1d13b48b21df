"""Device-resident state vector driven through the C ABI (include/tcb200.h).

PyTorch is used for exactly three things here: owning device memory (caching allocator),
naming the CUDA stream, and host<->device copies of small buffers.  No torch op ever touches
O(2^n) data; every such step is a tcb200 kernel.  There is no CPU path: constructing a
:class:`DeviceState` without a CUDA device raises."""

from __future__ import annotations

import os
from ctypes import c_void_p
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .fusion import KIND_DENSE, KIND_PERM, Block, plan_passes, plan_regtiles, tile_hi_fixpoint  # noqa: F401

_TORCH_C = {"complex64": torch.complex64, "complex128": torch.complex128}
_DT = {"complex64": _lib.C64, "complex128": _lib.C128}

# counters the benchmark reads (bytes are algorithmic: one read + one write per pass)
STATS = {"apply_launches": 0, "apply_bytes": 0, "expect_launches": 0, "sample_launches": 0,
         "gate_pass_rounds": 0, "gate_pass_fma_per_amp": 0.0, "gate_pass_free_gates": 0, "gate_pass_conflict_rounds": 0}


def reset_stats() -> None:
    for k in STATS:
        STATS[k] = 0


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise _lib.EngineError(
            "tensorcircuit_b200 needs a CUDA device (sm_100a); there is no CPU execution path"
        )


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor) -> c_void_p:
    return c_void_p(t.data_ptr())


def plan_expect_groups(flips: Sequence[int], nbits: int, tile_bits: int, max_hi: int = 5,
                       max_terms: int = _lib.MAX_TERMS) -> List[Tuple[List[int], List[int]]]:
    """Group Pauli terms into launches: each group holds <= max_terms terms whose flip bits fit
    one tile geometry (<= max_hi gathered bits).  Returns [(term ids, union of flip bits)]."""
    groups: List[Tuple[List[int], List[int]]] = []
    for t in range(len(flips)):
        fb = [b for b in range(nbits) if (int(flips[t]) >> b) & 1]
        if len(tile_hi_fixpoint(fb, tile_bits, nbits)) > max_hi:
            raise _lib.EngineError("Pauli string flips more than %d bits above the tile" % max_hi)
        placed = False
        for ids, union in groups:
            if len(ids) >= max_terms:
                continue
            u = sorted(set(union) | set(fb))
            if len(tile_hi_fixpoint(u, tile_bits, nbits)) <= max_hi:
                ids.append(t)
                union[:] = u
                placed = True
                break
        if not placed:
            groups.append(([t], sorted(fb)))
    return groups


def plan_single_flip_groups(flip_bits: Sequence[int], nbits: int, tile_bits: int, max_hi: int = 9, max_bits: int = 12,
                            per_bit: int = 1) -> List[Tuple[List[int], List[int]]]:
    """Launch plan of the single-flip kernel: each launch covers <= max_bits distinct flip bits that
    fit one tile ({low bits} U <= max_hi gathered bits), per_bit strings each.  High bits go
    first (they need a gathered slot; the low bits are inside every tile).  Returns
    [(term ids, gathered bits)]."""
    by_bit: dict = {}
    for i, b in enumerate(flip_bits):
        by_bit.setdefault(int(b), []).append(i)
    groups: List[Tuple[List[int], List[int]]] = []
    while by_bit:
        bits = sorted(by_bit, reverse=True)
        best: Optional[Tuple[int, int, List[int], List[int]]] = None
        for h in range(0, max_hi + 1):
            lrow = tile_bits - h
            if nbits <= tile_bits and h:
                break
            high = [b for b in bits if b >= lrow][:h]
            if len(high) < h:
                continue
            low = [b for b in bits if b < lrow]
            covered = high + low  # gathered bits first: the low bits are inside every tile
            score = (min(max_bits, len(covered)), h)
            if best is None or score > best[:2]:
                best = (score[0], h, sorted(high), covered[:max_bits])
        assert best is not None
        hi, chosen = best[2], best[3]
        # bits between lrow and the lowest gathered bit cannot be reached unless gathered
        ids: List[int] = []
        for b in chosen:
            take = by_bit[b][:per_bit]
            ids += take
            by_bit[b] = by_bit[b][per_bit:]
            if not by_bit[b]:
                del by_bit[b]
        if not ids:
            raise _lib.EngineError("single-flip planner made no progress")
        groups.append((ids, hi))
    return groups


class DeviceCOO:
    """A sparse operator resident on the device: int64 row / column amplitude indices (reference basis
    order) and complex128 values.  Built once (``backend.coo_sparse_matrix_from_numpy``,
    ``quantum.PauliStringSum2COO`` for a generic matrix) and reused over the iterations of a loop."""

    def __init__(self, rows: Any, cols: Any, vals: Any, dim: int, device: Any = None):
        require_cuda()
        r = np.ascontiguousarray(np.asarray(rows, dtype=np.int64).reshape(-1))
        c = np.ascontiguousarray(np.asarray(cols, dtype=np.int64).reshape(-1))
        v = np.ascontiguousarray(np.asarray(vals, dtype=np.complex128).reshape(-1))
        if not (r.size == c.size == v.size):
            raise ValueError("rows, cols and values must have the same length")
        if r.size and (min(r.min(), c.min()) < 0 or max(r.max(), c.max()) >= dim):
            raise ValueError("sparse operator index outside [0, %d)" % dim)
        dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.dim, self.nnz, self.shape = int(dim), int(r.size), (int(dim), int(dim))
        self.rows = torch.from_numpy(r).to(dev)
        self.cols = torch.from_numpy(c).to(dev)
        self.vals = torch.from_numpy(v).to(dev)

    def csr(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(indptr, indices, values) of the same operator in CSR form on the device, built once (host sort of the
        stored elements): what ``lambda = H psi`` of the adjoint sweep walks row by row"""
        if getattr(self, "_csr", None) is None:
            import scipy.sparse as sp

            m = sp.coo_matrix((self.vals.cpu().numpy(), (self.rows.cpu().numpy(), self.cols.cpu().numpy())), shape=self.shape).tocsr()
            m.sum_duplicates()
            herm = abs(m - m.getH())
            self.hermitian = bool(herm.nnz == 0 or herm.max() <= 1e-10 * max(1.0, abs(m).max()))
            dev = self.rows.device
            self._csr = (torch.from_numpy(m.indptr.astype(np.int64)).to(dev), torch.from_numpy(m.indices.astype(np.int64)).to(dev),
                         torch.from_numpy(m.data.astype(np.complex128)).to(dev))
        return self._csr

    @classmethod
    def from_scipy(cls, m: Any, device: Any = None) -> "DeviceCOO":
        m = m.tocoo()
        if m.shape[0] != m.shape[1]:
            raise ValueError("operator must be square")
        return cls(m.row, m.col, m.data, m.shape[0], device)


class DeviceState:
    use_single_flip = os.environ.get("TCB200_SINGLE_FLIP", "1") != "0"

    def __init__(self, nbits: int, dtype: str = "complex64", batch: int = 1, device: Any = None,
                 buffer: Optional[torch.Tensor] = None):
        require_cuda()
        if dtype not in _DT:
            raise ValueError(f"Unsupported data type: {dtype}")
        self.nbits = int(nbits)
        self.dtype = dtype
        self.dt = _DT[dtype]
        self.batch = int(batch)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if buffer is not None:
            assert buffer.is_cuda and buffer.dtype == _TORCH_C[dtype] and buffer.numel() == self.batch << self.nbits
            self.buf = buffer.view(self.batch, -1)
        else:
            self.buf = torch.empty((self.batch, 1 << self.nbits), dtype=_TORCH_C[dtype], device=self.device)
        self._ws: Optional[torch.Tensor] = None

    # -- helpers ----------------------------------------------------------------------------
    @property
    def amp_bytes(self) -> int:
        return 8 if self.dtype == "complex64" else 16

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def _dev_matrix(self, m: np.ndarray) -> torch.Tensor:
        h = np.ascontiguousarray(m.astype(np.complex64 if self.dtype == "complex64" else np.complex128))
        return torch.from_numpy(h).to(self.device, non_blocking=False)

    # -- initial state ------------------------------------------------------------------------
    def init_zero(self) -> None:
        check(lib.tcb200_init_zero(_ptr(self.buf), self.nbits, self.dt, self.batch, _stream()))

    def load(self, src: Any) -> None:
        """Copy an initial state (host array / torch tensor of 2^nbits entries) into every batch row."""
        if isinstance(src, torch.Tensor):
            s = src.reshape(-1).to(self.device, dtype=torch.complex128)
        else:
            s = torch.from_numpy(np.ascontiguousarray(np.asarray(src).reshape(-1).astype(np.complex128))).to(self.device)
        if s.numel() != 1 << self.nbits:
            raise ValueError("initial state has %d entries, expected %d" % (s.numel(), 1 << self.nbits))
        for b in range(self.batch):
            check(lib.tcb200_load_c128(_ptr(self.buf[b]), self.nbits, self.dt, _ptr(s), _stream()))
        torch.cuda.current_stream().synchronize()  # `s` may be freed by the allocator afterwards

    # -- gates ------------------------------------------------------------------------------
    def apply_block(self, blk: Block) -> None:
        k = len(blk.bits)
        bits = np.asarray(blk.bits, dtype=np.int32)
        if blk.batched:
            if blk.matrix.shape[0] != self.batch:
                raise ValueError("batched block of size %d on a state of batch %d" % (blk.matrix.shape[0], self.batch))
            md = self._dev_matrix(blk.matrix)
            check(lib.tcb200_apply_dense_batched(_ptr(self.buf), self.nbits, self.dt, k, _lib.iptr(bits), _ptr(md), self.batch, _stream()))
        else:
            m = np.ascontiguousarray(blk.matrix, dtype=np.complex128)
            check(lib.tcb200_apply_dense(_ptr(self.buf), self.nbits, self.dt, k, _lib.iptr(bits), _lib.dptr(m.view(np.float64)), self.batch, _stream()))
        STATS["apply_launches"] += 1
        STATS["apply_bytes"] += 2 * self.amp_bytes * (self.batch << self.nbits)

    def apply_blocks(self, blocks: Sequence[Block]) -> None:
        for b in blocks:
            self.apply_block(b)

    # widest set of gathered high bits a staged pass may use.  64 KiB tile, 8 gathered bits: the 5
    # low bits stay contiguous (256-byte rows for complex64).  Measured on the config-4 recipe:
    # n = 32: 60 / 47 / 44 / 41 passes and 1380 / 1247 / 1261 / 1212 ms for max_hi = 5 / 6 / 7 / 8;
    # n = 33: 2575 -> 2475 ms from 6 to 7 (profiles/README.md) -- fewer, fuller passes win although
    # the shorter rows cost some HBM efficiency per pass.
    pass_max_hi = int(os.environ.get("TCB200_PASS_MAX_HI", "8"))

    # The structure-aware gate pass (tcb200_apply_gate_pass) is the production path; the older
    # dense multi-block pass (cpass) stays reachable with TCB200_GATE_PASS=0 and serves batched
    # (vmap) matrices and states below 4 qubits.
    use_gate_pass = os.environ.get("TCB200_GATE_PASS", "1") != "0"
    gate_pass_max_ops = int(os.environ.get("TCB200_GATE_PASS_MAX_OPS", "200"))
    # gathered bits of a gate pass: 9 (128-byte rows for complex64) only in the production shape
    # (state larger than one 64 KiB tile), 8 otherwise
    gate_pass_max_hi = int(os.environ.get("TCB200_GATE_PASS_MAX_HI", "9"))

    def apply_planned(self, blocks: Sequence[Block]) -> int:
        """Run ``blocks`` as staged passes (fusion.plan_passes): each pass is one HBM read + write
        of the state however many blocks it holds.  Returns the number of launches."""
        if not blocks:
            return 0
        T = lib.tcb200_pass_tile_bits(self.dt)
        if self.use_gate_pass and self.nbits >= 4:
            return self._apply_gate_planned(blocks, T)
        mat_elems = (12 * 1024) // self.amp_bytes
        passes = plan_passes([b.bits for b in blocks], self.nbits, T, max_hi=self.pass_max_hi,
                             max_ops=_lib.MAX_PASS_OPS, max_mat_elems=mat_elems, max_pass_k=_lib.MAX_PASS_K)
        nlaunch = 0
        for p in passes:
            blks = [blocks[i] for i in p.block_ids]
            if len(blks) == 1:
                self.apply_block(blks[0])
                nlaunch += 1
            elif any(b.batched for b in blks):
                self.apply_pass(blks, p.tile_hi)
                nlaunch += 1
            elif self.use_regtiles and all(len(b.bits) <= 2 for b in blks):
                nlaunch += self.apply_rpass_host(blocks, p.block_ids, p.tile_hi)
            else:
                self.apply_pass_host(blks, p.tile_hi)
                nlaunch += 1
        return nlaunch

    def _apply_gate_planned(self, blocks: Sequence[Block], T: int) -> int:
        """Pass plan for the gate pass: permutation blocks are free (weight 0, no parameter-bank
        space), diagonal blocks take a 16-entry table, dense blocks wider than 3 bits run alone
        through the single-block kernel."""
        cost = [0 if b.kind == KIND_PERM else (4 ** len(b.bits) if b.kind == KIND_DENSE else 16) for b in blocks]
        weight = [0.0 if b.kind == KIND_PERM else 1.0 for b in blocks]
        max_hi = min(self.gate_pass_max_hi, 9 if (self.nbits > T and T - (1 if self.amp_bytes == 8 else 0) == 12) else 8)
        passes = plan_passes([b.bits for b in blocks], self.nbits, T, max_hi=max_hi, max_ops=self.gate_pass_max_ops,
                             max_mat_elems=_lib.GATE_PASS_MAT_ELEMS, max_pass_k=4, block_cost=cost, block_weight=weight,
                             block_diag=[b.kind == "diag" for b in blocks])
        nlaunch = 0
        for p in passes:
            blks = [blocks[i] for i in p.block_ids]
            if len(blks) == 1 and (len(blks[0].bits) > 3 and blks[0].kind == KIND_DENSE or len(blks[0].bits) > 4):
                self.apply_block(blks[0])
                nlaunch += 1
            else:
                nlaunch += self.apply_gate_pass(blks, p.tile_hi)
        return nlaunch

    def apply_gate_pass(self, blocks: Sequence[Block], tile_hi: Sequence[int]) -> int:
        """One structure-aware staged pass (include/tcb200.h: tcb200_apply_gate_pass).  A run that
        exceeds one launch's capacity is split in two on the same tile."""
        if any(len(b.bits) > 3 and b.kind == KIND_DENSE for b in blocks):
            # a dense block wider than the register tile handles: it runs alone, in order
            n = 0
            run: List[Block] = []
            for b in blocks:
                if len(b.bits) > 3 and b.kind == KIND_DENSE:
                    if run:
                        n += self.apply_gate_pass(run, tile_hi)
                        run = []
                    self.apply_block(b)
                    n += 1
                else:
                    run.append(b)
            if run:
                n += self.apply_gate_pass(run, tile_hi)
            return n
        ks = np.asarray([len(b.bits) for b in blocks], dtype=np.int32)
        bits = np.asarray([x for b in blocks for x in b.bits], dtype=np.int32)
        mats = np.ascontiguousarray(np.concatenate([np.asarray(b.matrix, dtype=np.complex128).reshape(-1) for b in blocks]))
        hi = np.asarray(list(tile_hi) if len(tile_hi) else [0], dtype=np.int32)
        info = np.zeros(8, dtype=np.float64)
        if any(b.batched for b in blocks):
            for b in blocks:
                if b.batched and b.matrix.shape[0] != self.batch:
                    raise ValueError("batched block of size %d on a state of batch %d" % (b.matrix.shape[0], self.batch))
            flags = np.asarray([1 if b.batched else 0 for b in blocks], dtype=np.int32)
            rc = self._gate_pass_call_batched(len(blocks), ks, bits, mats, flags, len(tile_hi), hi, info)
        else:
            rc = self._gate_pass_call(len(blocks), ks, bits, mats, len(tile_hi), hi, info)
        if rc == _lib.ERR_CAPACITY and len(blocks) > 1:
            h = len(blocks) // 2
            return self.apply_gate_pass(blocks[:h], tile_hi) + self.apply_gate_pass(blocks[h:], tile_hi)
        check(rc)
        STATS["apply_launches"] += 1
        STATS["apply_bytes"] += 2 * self.amp_bytes * (self.batch << self.nbits)
        STATS["gate_pass_rounds"] += int(info[0])
        STATS["gate_pass_free_gates"] += int(info[1])
        STATS["gate_pass_conflict_rounds"] += int(info[4])
        STATS["gate_pass_fma_per_amp"] += float(info[6])
        return 1

    def _gate_pass_call(self, nops: int, ks: np.ndarray, bits: np.ndarray, mats: np.ndarray, n_hi: int, hi: np.ndarray, info: np.ndarray) -> int:
        return lib.tcb200_apply_gate_pass(_ptr(self.buf), self.nbits, self.dt, nops, _lib.iptr(ks), _lib.iptr(bits), _lib.dptr(mats.view(np.float64)),
                                          n_hi, _lib.iptr(hi), self.batch, _lib.dptr(info), _stream())

    def _gate_pass_call_batched(self, nops: int, ks: np.ndarray, bits: np.ndarray, mats: np.ndarray, flags: np.ndarray, n_hi: int,
                                hi: np.ndarray, info: np.ndarray) -> int:
        need = int(lib.tcb200_gate_pass_batched_workspace_bytes(self.dt, self.batch))
        if getattr(self, "_gp_ws", None) is None or self._gp_ws.numel() < need:
            self._gp_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return lib.tcb200_apply_gate_pass_batched(_ptr(self.buf), self.nbits, self.dt, nops, _lib.iptr(ks), _lib.iptr(bits), _lib.dptr(mats.view(np.float64)),
                                                  _lib.iptr(flags), n_hi, _lib.iptr(hi), self.batch, _ptr(self._gp_ws), self._gp_ws.numel(),
                                                  _lib.dptr(info), _stream())

    # Register tiles (several gates per shared-memory round trip) pay off when many gates pile up
    # on the same <= 4 qubits inside a pass (nearest-neighbour ladders); on the random-matching
    # benchmark circuit a pass holds ~2 gates per tile and the generic dispatch costs more than it
    # saves (measured: profiles/README.md), so the plain multi-block pass is the default.
    use_regtiles = os.environ.get("TCB200_REGTILES", "0") != "0"

    # 2 = pair mode: two disjoint 2-bit gates per 4-bit register tile (the specialised path)
    regtile_max_gates = int(os.environ.get("TCB200_REGTILE_GATES", "2"))

    def apply_rpass_host(self, blocks: Sequence[Block], ids: Sequence[int], tile_hi: Sequence[int]) -> int:
        """One staged pass whose blocks are clustered into register tiles (fusion.plan_regtiles)."""
        kt_max = 4 if self.dtype == "complex64" else 3
        tiles = plan_regtiles([b.bits for b in blocks], list(ids), max_bits=kt_max, max_gates=self.regtile_max_gates)
        hi = np.asarray(list(tile_hi) if len(tile_hi) else [0], dtype=np.int32)
        launches = 0
        for c0 in range(0, len(tiles), _lib.MAX_PASS_OPS):
            chunk = tiles[c0 : c0 + _lib.MAX_PASS_OPS]
            rt_k = np.asarray([len(t.bits) for t in chunk], dtype=np.int32)
            rt_bits = np.asarray([x for t in chunk for x in t.bits], dtype=np.int32)
            rt_nsub = np.asarray([len(t.block_ids) for t in chunk], dtype=np.int32)
            sub_k = np.asarray([len(blocks[i].bits) for t in chunk for i in t.block_ids], dtype=np.int32)
            sub_bits = np.asarray([x for t in chunk for i in t.block_ids for x in blocks[i].bits], dtype=np.int32)
            mats = np.ascontiguousarray(np.concatenate([np.asarray(blocks[i].matrix, dtype=np.complex128).reshape(-1) for t in chunk for i in t.block_ids]))
            check(lib.tcb200_apply_rpass_host(_ptr(self.buf), self.nbits, self.dt, len(chunk), _lib.iptr(rt_k), _lib.iptr(rt_bits), _lib.iptr(rt_nsub),
                                              _lib.iptr(sub_k), _lib.iptr(sub_bits), _lib.dptr(mats.view(np.float64)), len(tile_hi), _lib.iptr(hi),
                                              self.batch, _stream()))
            STATS["apply_launches"] += 1
            STATS["apply_bytes"] += 2 * self.amp_bytes * (self.batch << self.nbits)
            launches += 1
        return launches

    def apply_pass_host(self, blocks: Sequence[Block], tile_hi: Sequence[int]) -> None:
        """Several shared-matrix blocks in one staged pass, matrices in the constant bank."""
        ks = np.asarray([len(b.bits) for b in blocks], dtype=np.int32)
        bits = np.asarray([x for b in blocks for x in b.bits], dtype=np.int32)
        mats = np.ascontiguousarray(np.concatenate([np.asarray(b.matrix, dtype=np.complex128).reshape(-1) for b in blocks]))
        hi = np.asarray(list(tile_hi) if len(tile_hi) else [0], dtype=np.int32)
        check(lib.tcb200_apply_pass_host(_ptr(self.buf), self.nbits, self.dt, len(blocks), _lib.iptr(ks), _lib.iptr(bits),
                                         _lib.dptr(mats.view(np.float64)), len(tile_hi), _lib.iptr(hi), self.batch, _stream()))
        STATS["apply_launches"] += 1
        STATS["apply_bytes"] += 2 * self.amp_bytes * (self.batch << self.nbits)

    def apply_pass(self, blocks: Sequence[Block], tile_hi: Sequence[int]) -> None:
        """Several blocks inside one staged tile pass (one HBM read + write)."""
        ks = np.asarray([len(b.bits) for b in blocks], dtype=np.int32)
        bits = np.asarray([x for b in blocks for x in b.bits], dtype=np.int32)
        batched = any(b.batched for b in blocks)
        mats = []
        for b in blocks:
            m = b.matrix
            if batched and m.ndim == 2:
                m = np.broadcast_to(m, (self.batch,) + m.shape)
            mats.append(np.ascontiguousarray(m).reshape(-1))
        md = self._dev_matrix(np.concatenate(mats))
        hi = np.asarray(list(tile_hi) if len(tile_hi) else [0], dtype=np.int32)
        check(lib.tcb200_apply_pass(_ptr(self.buf), self.nbits, self.dt, len(blocks), _lib.iptr(ks), _lib.iptr(bits), _ptr(md),
                                    self.batch if batched else 1, len(tile_hi), _lib.iptr(hi), self.batch, _stream()))
        STATS["apply_launches"] += 1
        STATS["apply_bytes"] += 2 * self.amp_bytes * (self.batch << self.nbits)

    # -- reductions ---------------------------------------------------------------------------
    def norm2(self) -> np.ndarray:
        ws = self._workspace(lib.tcb200_reduce_workspace_bytes(self.nbits, self.batch))
        out = torch.empty(self.batch, dtype=torch.float64, device=self.device)
        check(lib.tcb200_norm2(_ptr(self.buf), self.nbits, self.dt, self.batch, _ptr(out), _ptr(ws), ws.numel(), _stream()))
        return out.cpu().numpy()

    def probability_state(self) -> "DeviceState":
        """a new state-shaped buffer holding (|psi_e|^2, 0): the probabilities as something the
        apply kernels can transform (readout error, basecircuit.py:760-803)"""
        if self.batch != 1:
            raise _lib.EngineError("probability_state on a batched state is not supported")
        r = type(self)(self.nbits, self.dtype)
        check(lib.tcb200_probability_state(_ptr(self.buf), _ptr(r.buf), self.nbits, self.dt, 0, _stream()))
        return r

    def sqrt_real_inplace(self) -> None:
        """(p, *) -> (sqrt(max(p, 0)), 0): afterwards |amplitude|^2 is the distribution p"""
        check(lib.tcb200_probability_state(_ptr(self.buf), _ptr(self.buf), self.nbits, self.dt, 1, _stream()))

    def masked_norm2(self, mask: int, value: int) -> float:
        """sum of |psi_e|^2 over the amplitudes with (e & mask) == value (batch 1): the mass of a
        partial measurement record, basecircuit.py:359-443"""
        if self.batch != 1:
            raise _lib.EngineError("masked_norm2 on a batched state is not supported")
        ws = self._workspace(lib.tcb200_masked_norm2_workspace_bytes())
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        check(lib.tcb200_masked_norm2(_ptr(self.buf), self.nbits, self.dt, int(mask), int(value), _ptr(out), _ptr(ws), ws.numel(), _stream()))
        return float(out.cpu().numpy()[0])

    def probability(self) -> torch.Tensor:
        rd = torch.float32 if self.dtype == "complex64" else torch.float64
        p = torch.empty((self.batch, 1 << self.nbits), dtype=rd, device=self.device)
        check(lib.tcb200_probability(_ptr(self.buf), self.nbits, self.dt, _ptr(p), self.batch, _stream()))
        return p

    def expectation_terms(self, flips: Sequence[int], signs: Sequence[int], nys: Sequence[int]) -> np.ndarray:
        """<P_t> for every term (amplitude-bit masks) and batch element -> complex [batch, nterms].

        Terms are grouped so that each launch covers up to MAX_TERMS strings whose flip masks
        fit one tile geometry; each launch reads the state once."""
        nt = len(flips)
        out_all = np.zeros((self.batch, nt), dtype=np.complex128)
        if nt == 0:
            return out_all
        T = lib.tcb200_expect_tile_bits(self.dt)
        outs = []
        rest = list(range(nt))
        if self.nbits >= lib.tcb200_expect_z_min_bits(self.dt):
            # diagonal strings (no X / Y): 32 (16 for complex128) per streaming read of the state
            zids = [t for t in rest if int(flips[t]) == 0]
            rest = [t for t in rest if int(flips[t]) != 0]
            zt = lib.tcb200_expect_z_max_terms(self.dt)
            wsz = self._workspace(lib.tcb200_expect_z_workspace_bytes(self.nbits, self.batch))
            for c0 in range(0, len(zids), zt):
                ids = zids[c0 : c0 + zt]
                s = np.asarray([int(signs[t]) for t in ids], dtype=np.uint64)
                out = torch.empty((self.batch, len(ids), 2), dtype=torch.float64, device=self.device)
                check(lib.tcb200_expect_z(_ptr(self.buf), self.nbits, self.dt, len(ids), _lib.u64ptr(s), _ptr(out), self.batch,
                                          _ptr(wsz), wsz.numel(), _stream()))
                STATS["expect_launches"] += 1
                outs.append((ids, out))
        # single-flip strings (one X or Y, any Z's): 12 per read through the register-pair kernel
        Tp = lib.tcb200_pass_tile_bits(self.dt)
        if self.use_single_flip and self.nbits > Tp and Tp - (1 if self.amp_bytes == 8 else 0) == 12:
            sf = [t for t in rest if bin(int(flips[t])).count("1") == 1 and int(nys[t]) <= 1]
            if sf:
                rest = [t for t in rest if t not in set(sf)]
                wsx = self._workspace(max(lib.tcb200_expect_single_flip_workspace_bytes(self.nbits, self.batch),
                                          lib.tcb200_expect_workspace_bytes(self.nbits, self.batch),
                                          lib.tcb200_expect_z_workspace_bytes(self.nbits, self.batch) if self.nbits >= lib.tcb200_expect_z_min_bits(self.dt) else 0))
                for ids, hi in plan_single_flip_groups([int(flips[t]).bit_length() - 1 for t in sf], self.nbits, Tp):
                    tid = [sf[i] for i in ids]
                    fb = np.asarray([int(flips[t]).bit_length() - 1 for t in tid], dtype=np.int32)
                    s = np.asarray([int(signs[t]) for t in tid], dtype=np.uint64)
                    ny = np.asarray([int(nys[t]) for t in tid], dtype=np.int32)
                    hia = np.asarray(hi if hi else [0], dtype=np.int32)
                    out = torch.empty((self.batch, len(tid), 2), dtype=torch.float64, device=self.device)
                    check(lib.tcb200_expect_single_flip(_ptr(self.buf), self.nbits, self.dt, len(tid), _lib.iptr(fb), _lib.u64ptr(s), _lib.iptr(ny),
                                                        len(hi), _lib.iptr(hia), _ptr(out), self.batch, _ptr(wsx), wsx.numel(), _stream()))
                    STATS["expect_launches"] += 1
                    outs.append((tid, out))
        sub = plan_expect_groups([flips[t] for t in rest], self.nbits, T)
        groups = [([rest[i] for i in ids], union) for ids, union in sub]
        ws = self._workspace(lib.tcb200_expect_workspace_bytes(self.nbits, self.batch))
        for ids, union in groups:
            hi = tile_hi_fixpoint(union, T, self.nbits)
            f = np.asarray([int(flips[t]) for t in ids], dtype=np.uint64)
            s = np.asarray([int(signs[t]) for t in ids], dtype=np.uint64)
            ny = np.asarray([int(nys[t]) for t in ids], dtype=np.int32)
            hia = np.asarray(hi if hi else [0], dtype=np.int32)
            out = torch.empty((self.batch, len(ids), 2), dtype=torch.float64, device=self.device)
            check(lib.tcb200_expect_pauli(_ptr(self.buf), self.nbits, self.dt, len(ids), _lib.u64ptr(f), _lib.u64ptr(s), _lib.iptr(ny),
                                          len(hi), _lib.iptr(hia), _ptr(out), self.batch, _ptr(ws), ws.numel(), _stream()))
            STATS["expect_launches"] += 1
            outs.append((ids, out))
        for ids, out in outs:
            o = out.cpu().numpy()
            out_all[:, ids] = o[..., 0] + 1j * o[..., 1]
        return out_all

    # -- sparse operators and the adjoint-sweep primitives (csrc/sparse.cu) -------------------------
    def coo_expectation(self, op: "DeviceCOO") -> np.ndarray:
        """<psi| H |psi> per batch element for a device-resident COO operator: complex [batch]."""
        if op.dim != 1 << self.nbits:
            raise ValueError("operator of dimension %d on a state of %d qubits" % (op.dim, self.nbits))
        ws = self._workspace(lib.tcb200_coo_expectation_workspace_bytes(op.nnz, self.batch))
        out = torch.empty((self.batch, 2), dtype=torch.float64, device=self.device)
        check(lib.tcb200_coo_expectation(_ptr(self.buf), self.nbits, self.dt, op.nnz, _ptr(op.rows), _ptr(op.cols), _ptr(op.vals), _ptr(out),
                                         self.batch, _ptr(ws), ws.numel(), _stream()))
        STATS["expect_launches"] += 1
        o = out.cpu().numpy()
        return o[:, 0] + 1j * o[:, 1]

    def row_state(self, b: int) -> "DeviceState":
        """row ``b`` of a batched state as an unbatched state on the same memory (no copy)"""
        return type(self)(self.nbits, self.dtype, 1, self.device, buffer=self.buf[b])

    def inner(self, bra: "DeviceState", row: int = 0, bra_row: int = 0) -> complex:
        """<bra|self> through the transition-element kernel with the identity on bit 0"""
        assert bra.nbits == self.nbits and bra.dtype == self.dtype
        ks = np.asarray([1], dtype=np.int32)
        bits = np.asarray([0], dtype=np.int32)
        mats = np.ascontiguousarray(np.eye(2, dtype=np.complex128).reshape(-1))
        ws = self._workspace(lib.tcb200_transition_local_workspace_bytes(1, self.nbits, self.dt))
        out = torch.empty((1, 2), dtype=torch.float64, device=self.device)
        check(lib.tcb200_transition_local(_ptr(bra.buf[bra_row]), _ptr(self.buf[row]), self.nbits, self.dt, 1, _lib.iptr(ks), _lib.iptr(bits),
                                          _lib.dptr(mats.view(np.float64)), _ptr(out), _ptr(ws), ws.numel(), _stream()))
        STATS["expect_launches"] += 1
        o = out.cpu().numpy()
        return complex(o[0, 0], o[0, 1])

    def copy_row_from(self, row: int, src: "DeviceState", src_row: int = 0) -> None:
        """self[row] <- src[src_row] (device-to-device, same size and dtype)."""
        assert src.nbits == self.nbits and src.dtype == self.dtype
        nb = self.amp_bytes << self.nbits
        check(lib.tcb200_copy_rows(_ptr(self.buf[row]), nb, _ptr(src.buf[src_row]), nb, nb, 1, _stream()))

    def apply_pauli_sum_rows(self, src_row: int, dst_row: int, flips: Sequence[int], signs: Sequence[int], coef: Sequence[complex]) -> None:
        """self[dst_row] <- sum_t coef_t P_t self[src_row]  (coef_t already holds (-i)^ny_t)."""
        assert src_row != dst_row
        nt = len(flips)
        f = np.asarray([int(x) for x in flips], dtype=np.uint64)
        g = np.asarray([int(x) for x in signs], dtype=np.uint64)
        c = np.ascontiguousarray(np.asarray(coef, dtype=np.complex128))
        ws = self._workspace(lib.tcb200_apply_pauli_sum_workspace_bytes(nt))
        check(lib.tcb200_apply_pauli_sum(_ptr(self.buf[src_row]), _ptr(self.buf[dst_row]), self.nbits, self.dt, nt, _lib.u64ptr(f), _lib.u64ptr(g),
                                         _lib.dptr(c.view(np.float64)), 1, _ptr(ws), ws.numel(), _stream()))
        STATS["apply_launches"] += 1

    def apply_csr_rows(self, src_row: int, dst_row: int, op: "DeviceCOO", coef: complex = 1.0, accumulate: bool = False) -> None:
        """self[dst_row] (+)= coef * H self[src_row] for a device-resident sparse operator"""
        assert src_row != dst_row
        if op.dim != 1 << self.nbits:
            raise ValueError("operator of dimension %d on a state of %d qubits" % (op.dim, self.nbits))
        indptr, indices, vals = op.csr()
        c = complex(coef)
        check(lib.tcb200_csr_matvec(_ptr(self.buf[src_row]), _ptr(self.buf[dst_row]), self.nbits, self.dt, _ptr(indptr), _ptr(indices), _ptr(vals),
                                    c.real, c.imag, 1 if accumulate else 0, _stream()))
        STATS["apply_launches"] += 1

    def transition_local(self, bra_row: int, ket_row: int, ops: Sequence[Tuple[Sequence[int], np.ndarray]]) -> np.ndarray:
        """<self[bra_row]| G_j |self[ket_row]> for local operators (bits ascending, matrix index bit i <->
        bits[i]): complex [nops]; tcb200_transition_local_max_ops() operators per launch."""
        res = np.zeros(len(ops), dtype=np.complex128)
        per = lib.tcb200_transition_local_max_ops()
        for c0 in range(0, len(ops), per):
            chunk = ops[c0 : c0 + per]
            ks = np.asarray([len(b) for b, _ in chunk], dtype=np.int32)
            bits = np.asarray([x for b, _ in chunk for x in b], dtype=np.int32)
            mats = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.complex128).reshape(-1) for _, m in chunk]))
            ws = self._workspace(lib.tcb200_transition_local_workspace_bytes(len(chunk), self.nbits, self.dt))
            out = torch.empty((len(chunk), 2), dtype=torch.float64, device=self.device)
            check(lib.tcb200_transition_local(_ptr(self.buf[bra_row]), _ptr(self.buf[ket_row]), self.nbits, self.dt, len(chunk), _lib.iptr(ks),
                                              _lib.iptr(bits), _lib.dptr(mats.view(np.float64)), _ptr(out), _ptr(ws), ws.numel(), _stream()))
            STATS["expect_launches"] += 1
            o = out.cpu().numpy()
            res[c0 : c0 + len(chunk)] = o[:, 0] + 1j * o[:, 1]
        return res

    # -- sampler ------------------------------------------------------------------------------
    def sample(self, uniforms: Any, cdf_offset: float = 0.0, cdf_total: float = -1.0, return_total: bool = False) -> Any:
        if self.batch != 1:
            raise _lib.EngineError("sampling a batched state is not supported")
        if isinstance(uniforms, torch.Tensor):  # e.g. pinned host memory: async copy, no staging
            ud = uniforms.reshape(-1).to(self.device, dtype=torch.float64, non_blocking=True)
        else:
            u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(-1))
            ud = torch.from_numpy(u).to(self.device)
        shots = ud.shape[0]
        idx = torch.empty(max(shots, 1), dtype=torch.int64, device=self.device)
        tot = torch.empty(1, dtype=torch.float64, device=self.device)
        ws = self._workspace(lib.tcb200_sample_workspace_bytes(self.nbits))
        check(lib.tcb200_sample(_ptr(self.buf), self.nbits, self.dt, _ptr(ud), shots, _ptr(idx), _ptr(tot), float(cdf_offset), float(cdf_total),
                                _ptr(ws), ws.numel(), _stream()))
        STATS["sample_launches"] += 1
        res = idx[:shots].cpu().numpy()
        if return_total:
            return res, float(tot.cpu().item())
        return res
