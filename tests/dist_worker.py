"""Worker for the world_size>1 CPU tests (gloo): the distributed scheduler / remap / reductions of
tensorcircuit_b200.dist with the emulated local engine (tests/fake_state.py), checked against
the oracle.  Launched by tests/test_dist_gloo.py through torch.distributed.run."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import tc_oracle as orc  # noqa: E402
from tensorcircuit_b200 import recipes  # noqa: E402
from tensorcircuit_b200.circuit import Circuit  # noqa: E402
from tensorcircuit_b200.dist import DistState  # noqa: E402
from tensorcircuit_b200.fusion import fuse  # noqa: E402
from tests.fake_state import FakeState  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 10
    for mode in ("double", "chunked"):
        for dtype, tol in (("complex64", 2e-5), ("complex128", 1e-11)):
            ops = recipes.random_circuit(n, 5, seed=5) + recipes.hea_circuit(n, np.random.default_rng(1).uniform(0, 6, size=[2, 2, n]))
            c = recipes.build(Circuit(n), ops)
            blocks = fuse(c._ops, n, kmax=4)
            ds = DistState(n, dtype, state_factory=FakeState, double_buffer=(mode == "double"), staging_bytes=1 << 9)
            ds.init_zero()
            ds.run(blocks)
            o = orc.run_gatelist(n, ops)
            got = ds.gather_state()
            err = np.linalg.norm(got - o.state()) / np.linalg.norm(o.state())
            assert err < tol, (mode, dtype, err)
            assert ds.stats["remaps"] >= 1
            assert abs(ds.norm2() - 1.0) < 1e-4
            # expectations: TFIM strings + random strings (X/Y on every qubit incl. global ones)
            terms = recipes.tfim_terms(n)
            pss = [ps for _, ps in terms] + [list(r) for r in np.random.default_rng(0).integers(0, 4, size=[4, n])]
            fl, sg, ny, want = [], [], [], []
            for ps in pss:
                x, y, z = orc.resolve_ps(n, ps=ps)
                f, s, k = orc.pauli_masks(n, x, y, z)
                # dense random strings may flip more bits than one tile gathers: keep sparse ones
                fl.append(f); sg.append(s); ny.append(k)
                want.append(o.expectation_ps(ps=ps))
            vals = ds.expectation_terms(fl, sg, ny)
            assert np.max(np.abs(vals - np.array(want))) < 20 * tol, (mode, dtype)
            # state unchanged (as a logical vector) by the remaps done for the expectations
            got2 = ds.gather_state()
            assert np.linalg.norm(got2 - o.state()) / np.linalg.norm(o.state()) < tol
            # mass of partial measurement records (masks over local AND currently-global bits)
            p_all = np.abs(o.state()) ** 2
            idx_all = np.arange(2**n)
            for mask, value in ((1 << (n - 1), 0), ((1 << (n - 1)) | 1, 1), (0b1011 << (n - 4), 0b1001 << (n - 4)), (0, 0)):
                want_m = float(np.sum(p_all[(idx_all & mask) == value]))
                assert abs(ds.masked_norm2(mask, value) - want_m) < 20 * tol, (mode, dtype, mask)
            # sampling: identical indices on every rank, equal to the oracle rule
            u = np.random.default_rng(4).random(3000)
            s_idx = ds.sample(u)
            cdf = orc.sample_cdf(np.abs(got2) ** 2)
            # the distributed CDF runs over *physical* order; compare distributions, not indices
            p = np.abs(got2) ** 2
            assert np.all(p[s_idx] > 0)
            for q in range(n):
                emp = np.mean(1 - 2 * ((s_idx >> (n - 1 - q)) & 1))
                assert abs(emp - o.expectation_ps(z=[q]).real) < 0.08, (q, emp)
            chk = torch.from_numpy(s_idx.copy())
            dist.broadcast(chk, 0)
            assert np.array_equal(chk.numpy(), s_idx)
            # logical order: the layout goes back to the identity (<= 2 remaps), the state is unchanged as a logical
            # vector, and the indices are those of the reference rule on the flat vector (circuit.py:915-935)
            assert ds.phys != list(range(n))
            remaps0 = ds.stats["remaps"]
            s_log = ds.sample(u, logical_order=True)
            assert ds.phys == list(range(n)) and ds.stats["remaps"] - remaps0 <= 2
            got3 = ds.gather_state()
            assert np.linalg.norm(got3 - o.state()) / np.linalg.norm(o.state()) < tol
            p3 = np.abs(got3.astype(np.complex128)) ** 2
            cdf3 = np.cumsum(p3)
            r3 = cdf3[-1] * (1 - u)
            want3 = np.minimum(np.searchsorted(cdf3, r3, side="left"), 2**n - 1)
            bad = np.nonzero(s_log != want3)[0]
            assert len(bad) < 10, (mode, dtype, len(bad))
            for i in bad:  # ties between adjacent CDF values only
                lo_ = min(int(s_log[i]), int(want3[i]))
                assert abs(cdf3[lo_] - r3[i]) <= 1e-9 * cdf3[-1], (i, s_log[i], want3[i])
            assert np.array_equal(ds.sample(u, logical_order=True), s_log)  # idempotent, no further movement
            # gates after the restored layout still run (the map is consistent)
            ds.run(blocks[:6])
            psi = o.state().astype(np.complex128)
            for blk in blocks[:6]:
                k = len(blk.qubits)
                t = np.tensordot(np.asarray(blk.matrix).reshape([2] * (2 * k)), psi.reshape([2] * n), axes=(list(range(k, 2 * k)), list(blk.qubits)))
                psi = np.moveaxis(t, list(range(k)), list(blk.qubits)).reshape(-1)
            got4 = ds.gather_state()
            assert np.linalg.norm(got4 - psi) / np.linalg.norm(psi) < 2 * tol, (mode, dtype)
    # vmap batch sharding (no data-path collective; one all-gather at the end)
    import tensorcircuit_b200 as tc

    tc.engine.DeviceState = FakeState
    B = 5

    def f(th):
        c = tc.Circuit(3)
        c.rx(0, theta=th[0])
        c.cnot(0, 1)
        c.ry(2, theta=th[1])
        return tc.backend.real(c.expectation_ps(z=[1]) + c.expectation_ps(x=[2]))

    th = np.random.default_rng(0).uniform(0, 3, size=(B, 2))
    got = tc.backend.vmap(f)(th)
    want = np.cos(th[:, 0]) + np.sin(th[:, 1])
    assert got.shape == (B,) and np.allclose(got, want, atol=1e-5)
    # public API: c.sample(status=...) on the sharded state in logical order == the single-process run, index by index
    # (circuit.py:915-935; complex128 so that the two pass plans agree far below the CDF spacing)
    from tensorcircuit_b200.dist import DistState as _DS

    tc.set_dtype("complex128")
    try:
        n2 = 9
        ops2 = recipes.random_circuit(n2, 4, seed=11)
        u2 = np.random.default_rng(9).random(500)
        tc.set_distributed(True, sample_order="logical")
        cd = recipes.build(tc.Circuit(n2), ops2)
        s_d = np.asarray(cd.sample(batch=500, allow_state=True, status=u2, format="sample_int"))
        assert cd._state.ds.stats["remaps"] >= 1
        tc.set_distributed(False)
        c1 = recipes.build(tc.Circuit(n2), ops2)
        s_1 = np.asarray(c1.sample(batch=500, allow_state=True, status=u2, format="sample_int"))
        assert int(np.sum(s_d != s_1)) <= 1, int(np.sum(s_d != s_1))
        # in place (physical order) the same uniforms give the same distribution but other indices
        tc.set_distributed(True, sample_order="physical")
        cp = recipes.build(tc.Circuit(n2), ops2)
        s_p = np.asarray(cp.sample(batch=500, allow_state=True, status=u2, format="sample_int"))
        assert s_p.shape == s_1.shape
    finally:
        tc.set_distributed(False)
        _DS.sample_order = "physical"
        tc.set_dtype("complex64")
    dist.barrier()
    if rank == 0:
        print("DIST_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
