"""CPU stand-in for ``engine.DeviceState`` -- TEST INFRASTRUCTURE ONLY.

Lets the host-side logic (Circuit recording, fusion, Pauli decomposition, expectation grouping,
vmap batching, sample formats) run in the GPU-less build container.  Gate blocks and
expectations execute the real kernel bodies through tests/emu (CPU emulation of the CUDA
code); the sampler uses the oracle's restatement of the reference rule.  The product never
imports this module: ``tensorcircuit_b200.engine.DeviceState`` raises without a CUDA device."""

import ctypes

import numpy as np
import torch

from oracle import tc_oracle as orc
from tensorcircuit_b200 import engine

from .test_emu_kernels import EMU_LIB, _build_emu


def _ip(a):
    return np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class FakeState:
    _lib = None

    def __init__(self, nbits, dtype="complex64", batch=1, device=None, buffer=None):
        if FakeState._lib is None:
            _build_emu()
            FakeState._lib = ctypes.CDLL(EMU_LIB)
            FakeState._lib.emu_last_error.restype = ctypes.c_char_p
        self.nbits, self.dtype, self.batch = int(nbits), dtype, int(batch)
        self.dt = 0 if dtype == "complex64" else 1
        self.np = np.zeros((self.batch, 1 << self.nbits), dtype=dtype) if buffer is None else buffer.reshape(self.batch, -1)
        self.buf = torch.from_numpy(self.np)  # shares memory
        self.device = torch.device("cpu")

    @property
    def amp_bytes(self):
        return 8 if self.dtype == "complex64" else 16

    def init_zero(self):
        self.np[:] = 0
        self.np[:, 0] = 1

    def load(self, src):
        if isinstance(src, torch.Tensor):
            src = src.cpu().numpy()
        s = np.asarray(src).reshape(-1)
        if s.size != 1 << self.nbits:
            raise ValueError("initial state has %d entries" % s.size)
        self.np[:] = s.astype(self.dtype)[None, :]

    def apply_block(self, blk):
        bits = list(blk.bits)
        for b in range(self.batch):
            m = blk.matrix[b] if blk.batched else blk.matrix
            m = np.ascontiguousarray(m, dtype=np.complex128)
            row = self.np[b]
            rc = self._lib.emu_apply_dense(row.ctypes.data_as(ctypes.c_void_p), self.nbits, self.dt, len(bits), _ip(bits),
                                           m.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
            assert rc == 0, self._lib.emu_last_error()
        engine.STATS["apply_launches"] += 1

    def apply_blocks(self, blocks):
        for b in blocks:
            self.apply_block(b)

    pass_max_hi = 6
    use_gate_pass = True
    gate_pass_max_ops = 200
    gate_pass_max_hi = 6
    apply_gate_pass = engine.DeviceState.apply_gate_pass
    _apply_gate_planned = engine.DeviceState._apply_gate_planned

    def _gate_pass_call(self, nops, ks, bits, mats, n_hi, hi, info):
        st = self.np[0].copy() if self.batch > 1 else None
        for r in range(self.batch):
            rc = self._lib.emu_apply_gate_pass(self.np[r].ctypes.data_as(ctypes.c_void_p), self.nbits, self.dt, nops, _ip(ks), _ip(bits),
                                               mats.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_hi, _ip(hi),
                                               info.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
            if rc:
                assert r == 0
                if rc != -4:
                    raise engine._lib.EngineError("emu gate pass: %s" % self._lib.emu_last_error().decode())
                return rc
        del st
        return 0

    def _gate_pass_call_batched(self, nops, ks, bits, mats, flags, n_hi, hi, info):
        rc = self._lib.emu_apply_gate_pass_batched(self.np.ctypes.data_as(ctypes.c_void_p), self.nbits, self.dt, nops, _ip(ks), _ip(bits),
                                                   mats.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _ip(flags), n_hi, _ip(hi),
                                                   self.batch, info.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        if rc and rc != -4:
            raise engine._lib.EngineError("emu gate pass: %s" % self._lib.emu_last_error().decode())
        return rc

    regtile_max_gates = 2  # pair mode, as the engine; tests override it to cover the generic path

    def apply_planned(self, blocks):
        """Same planner as DeviceState.apply_planned; passes run through the emulated pass kernel."""
        from tensorcircuit_b200.fusion import plan_passes

        if not blocks:
            return 0
        T = self._lib.emu_pass_tile_bits(self.dt)
        if self.use_gate_pass and self.nbits >= 4:
            # the engine's own planning code, with the launch replaced by the emulated kernel
            return self._apply_gate_planned(blocks, T)
        passes = plan_passes([b.bits for b in blocks], self.nbits, T, max_hi=self.pass_max_hi, max_ops=16,
                             max_mat_elems=(12 * 1024) // self.amp_bytes, max_pass_k=4)
        assert sorted(i for p in passes for i in p.block_ids) == list(range(len(blocks)))
        for p in passes:
            blks = [blocks[i] for i in p.block_ids]
            if len(blks) == 1 or any(b.batched for b in blks):
                for b in blks:
                    self.apply_block(b)
                continue
            if all(len(b.bits) <= 2 for b in blks):  # register-tile pass, as DeviceState.apply_rpass_host
                from tensorcircuit_b200.fusion import plan_regtiles

                tiles = plan_regtiles([b.bits for b in blocks], list(p.block_ids), max_bits=4 if self.dt == 0 else 3,
                                      max_gates=self.regtile_max_gates)
                assert sorted(i for t in tiles for i in t.block_ids) == sorted(p.block_ids)
                rt_k = [len(t.bits) for t in tiles]
                rt_bits = [x for t in tiles for x in t.bits]
                rt_nsub = [len(t.block_ids) for t in tiles]
                sub_k = [len(blocks[i].bits) for t in tiles for i in t.block_ids]
                sub_bits = [x for t in tiles for i in t.block_ids for x in blocks[i].bits]
                mats = np.ascontiguousarray(np.concatenate([np.asarray(blocks[i].matrix, dtype=np.complex128).reshape(-1) for t in tiles for i in t.block_ids]))
                for r in range(self.batch):
                    rc = self._lib.emu_apply_rpass(self.np[r].ctypes.data_as(ctypes.c_void_p), self.nbits, self.dt, len(tiles), _ip(rt_k), _ip(rt_bits), _ip(rt_nsub),
                                                   _ip(sub_k), _ip(sub_bits), mats.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                   len(p.tile_hi), _ip(p.tile_hi if p.tile_hi else [0]))
                    assert rc == 0, self._lib.emu_last_error()
                engine.STATS["apply_launches"] += 1
                continue
            ks = [len(b.bits) for b in blks]
            bits = [x for b in blks for x in b.bits]
            mats = np.ascontiguousarray(np.concatenate([np.asarray(b.matrix, dtype=np.complex128).reshape(-1) for b in blks]))
            for r in range(self.batch):
                rc = self._lib.emu_apply_pass(self.np[r].ctypes.data_as(ctypes.c_void_p), self.nbits, self.dt, len(blks), _ip(ks), _ip(bits),
                                              mats.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(p.tile_hi), _ip(p.tile_hi if p.tile_hi else [0]))
                assert rc == 0, self._lib.emu_last_error()
            engine.STATS["apply_launches"] += 1
        return len(passes)

    def norm2(self):
        return np.sum(np.abs(self.np.astype(np.complex128)) ** 2, axis=1)

    def probability_state(self):
        r = type(self)(self.nbits, self.dtype)
        r.np[0] = (np.abs(self.np[0]) ** 2).astype(self.np.dtype)
        return r

    def sqrt_real_inplace(self):
        self.np[0] = np.sqrt(np.maximum(self.np[0].real, 0)).astype(self.np.dtype)

    def masked_norm2(self, mask, value):
        idx = np.arange(1 << self.nbits, dtype=np.uint64)
        sel = (idx & np.uint64(mask)) == np.uint64(value)
        return float(np.sum(np.abs(self.np[0].astype(np.complex128)[sel]) ** 2))

    def probability(self):
        rd = np.float32 if self.dtype == "complex64" else np.float64
        return torch.from_numpy((np.abs(self.np) ** 2).astype(rd))

    def expectation_terms(self, flips, signs, nys):
        nt = len(flips)
        out = np.zeros((self.batch, nt), dtype=np.complex128)
        T = self._lib.emu_expect_tile_bits(self.dt)
        for ids, union in engine.plan_expect_groups(flips, self.nbits, T):
            hi = engine.tile_hi_fixpoint(union, T, self.nbits)
            f = np.asarray([int(flips[t]) for t in ids], dtype=np.uint64)
            s = np.asarray([int(signs[t]) for t in ids], dtype=np.uint64)
            ny = np.asarray([int(nys[t]) for t in ids], dtype=np.int32)
            for b in range(self.batch):
                o = np.zeros(2 * len(ids))
                rc = self._lib.emu_expect(self.np[b].ctypes.data_as(ctypes.c_void_p), self.nbits, self.dt, len(ids),
                                          f.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), s.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                          _ip(ny), len(hi), _ip(hi if hi else [0]), o.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
                assert rc == 0, self._lib.emu_last_error()
                out[b, ids] = o[0::2] + 1j * o[1::2]
        return out

    def sample(self, uniforms, cdf_offset=0.0, cdf_total=-1.0, return_total=False):
        p = np.abs(self.np[0].astype(np.complex128)) ** 2
        if cdf_total < 0:
            r = orc.probability_sample(p, uniforms)
        else:  # shard of a distributed state: same contract as tcb200_sample
            cdf = np.cumsum(p)
            t = cdf_total * (1 - np.asarray(uniforms, dtype=np.float64)) - cdf_offset
            own = (t > 0) & (t <= cdf[-1])
            r = np.where(own, np.minimum(np.searchsorted(cdf, t, side="left"), p.size - 1), -1).astype(np.int64)
        return (r, float(p.sum())) if return_total else r

    # ---- sparse operators / adjoint-sweep primitives: numpy restatements of csrc/sparse.cu ----
    def coo_expectation(self, op):
        out = np.zeros(self.batch, dtype=np.complex128)
        for b in range(self.batch):
            psi = self.np[b].astype(np.complex128)
            out[b] = np.sum(np.conj(psi[op.rows]) * op.vals * psi[op.cols])
        return out

    def row_state(self, b):
        return type(self)(self.nbits, self.dtype, 1, buffer=self.np[b])

    def inner(self, bra, row=0, bra_row=0):
        return complex(np.vdot(bra.np[bra_row].astype(np.complex128), self.np[row].astype(np.complex128)))

    def copy_row_from(self, row, src, src_row=0):
        self.np[row] = src.np[src_row]

    def apply_pauli_sum_rows(self, src_row, dst_row, flips, signs, coef):
        r = np.arange(1 << self.nbits, dtype=np.uint64)
        src = self.np[src_row].astype(np.complex128)
        acc = np.zeros_like(src)
        for f, g, c in zip(flips, signs, coef):
            par = np.zeros(r.size, dtype=np.int64)
            m = r & np.uint64(int(g))
            for i in range(self.nbits):
                par ^= ((m >> np.uint64(i)) & np.uint64(1)).astype(np.int64)
            acc += c * (1 - 2 * par) * src[r ^ np.uint64(int(f))]
        self.np[dst_row] = acc.astype(self.np.dtype)

    def apply_csr_rows(self, src_row, dst_row, op, coef=1.0, accumulate=False):
        import scipy.sparse as sp

        m = sp.coo_matrix((op.vals, (op.rows, op.cols)), shape=op.shape).tocsr()
        op.hermitian = bool(abs(m - m.getH()).max() <= 1e-10) if m.nnz else True
        r = coef * (m @ self.np[src_row].astype(np.complex128))
        self.np[dst_row] = (self.np[dst_row].astype(np.complex128) + r if accumulate else r).astype(self.np.dtype)

    def transition_local(self, bra_row, ket_row, ops):
        n = self.nbits
        bra = self.np[bra_row].astype(np.complex128)
        ket = self.np[ket_row].astype(np.complex128)
        out = np.zeros(len(ops), dtype=np.complex128)
        for j, (bits, m) in enumerate(ops):
            k = len(bits)
            axes = [n - 1 - b for b in bits][::-1]  # matrix index big-endian = (bits[k-1], ..., bits[0])
            t = np.tensordot(np.asarray(m, dtype=np.complex128).reshape([2] * (2 * k)), ket.reshape([2] * n), axes=(list(range(k, 2 * k)), axes))
            t = np.moveaxis(t, list(range(k)), axes)
            out[j] = np.vdot(bra, t.reshape(-1))
        return out


class FakeCOO:
    """host stand-in for engine.DeviceCOO (same checks, numpy arrays)"""

    def __init__(self, rows, cols, vals, dim, device=None):
        self.rows = np.asarray(rows, dtype=np.int64).reshape(-1)
        self.cols = np.asarray(cols, dtype=np.int64).reshape(-1)
        self.vals = np.asarray(vals, dtype=np.complex128).reshape(-1)
        if self.rows.size and (min(self.rows.min(), self.cols.min()) < 0 or max(self.rows.max(), self.cols.max()) >= dim):
            raise ValueError("sparse operator index outside [0, %d)" % dim)
        self.dim, self.nnz, self.shape = int(dim), int(self.rows.size), (int(dim), int(dim))

    def csr(self):
        import scipy.sparse as sp

        m = sp.coo_matrix((self.vals, (self.rows, self.cols)), shape=self.shape).tocsr()
        self.hermitian = bool(abs(m - m.getH()).max() <= 1e-10) if m.nnz else True
        return m

    @classmethod
    def from_scipy(cls, m, device=None):
        m = m.tocoo()
        return cls(m.row, m.col, m.data, m.shape[0])
