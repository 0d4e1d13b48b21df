"""Multi-GPU worker (NCCL, real kernels): distributed state vs the oracle at n = 22, both exchange
modes, expectations with X/Y on global qubits, distributed sampler, and the public
``tc.set_distributed(True)`` path.  Launched by tests/test_dist_gpu.py through torchrun."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import tc_oracle as orc  # noqa: E402
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import recipes  # noqa: E402
from tensorcircuit_b200.dist import DistState  # noqa: E402
from tensorcircuit_b200.fusion import fuse  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 22
    ops = recipes.random_circuit(n, 4, seed=5)
    o = orc.run_gatelist(n, ops)
    ref = o.state()
    terms = recipes.tfim_terms(n)
    pss = [ps for _, ps in terms]
    want_e = np.array([o.expectation_ps(ps=ps) for ps in pss[:6] + pss[n : n + 6]])
    for mode in ("double", "chunked"):
        for kf in (2, 3):
            c = recipes.build(tc.Circuit(n), ops)
            blocks = fuse(c._ops, n, kmax=kf)
            ds = DistState(n, "complex64", double_buffer=(mode == "double"), staging_bytes=1 << 22)
            ds.init_zero()
            ds.run(blocks)
            got = ds.gather_state()
            err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            assert err < 1e-5, (mode, kf, err)
            assert ds.stats["remaps"] >= 1
            assert abs(ds.norm2() - 1.0) < 1e-5
            fl, sg, ny = [], [], []
            for ps in pss[:6] + pss[n : n + 6]:
                x, y, z = orc.resolve_ps(n, ps=ps)
                f, s, k = orc.pauli_masks(n, x, y, z)
                fl.append(f); sg.append(s); ny.append(k)
            vals = ds.expectation_terms(fl, sg, ny)
            assert np.max(np.abs(vals - want_e)) < 1e-5, (mode, kf)
            u = np.random.default_rng(4).random(50000)
            s_idx = ds.sample(u)
            p = np.abs(ref) ** 2
            assert np.all(p[s_idx] > 0)
            for q in range(n):
                emp = np.mean(1 - 2 * ((s_idx >> (n - 1 - q)) & 1))
                assert abs(emp - o.expectation_ps(z=[q]).real) < 0.03, (q, emp)
            # logical order: identity layout restored (<= 2 more remaps), then the single-GPU rule on the flat vector
            # (circuit.py:915-935); mismatches only where a uniform falls between two adjacent CDF values
            remaps0 = ds.stats["remaps"]
            s_log = ds.sample(u, logical_order=True)
            assert ds.phys == list(range(n)) and ds.stats["remaps"] - remaps0 <= 2
            dev = ds.gather_state().astype(np.complex128)
            assert np.linalg.norm(dev - ref) / np.linalg.norm(ref) < 1e-5
            cdf = np.cumsum(np.abs(dev) ** 2)
            r = cdf[-1] * (1 - u)
            want_s = np.minimum(np.searchsorted(cdf, r, side="left"), 2**n - 1)
            bad = np.nonzero(s_log != want_s)[0]
            assert len(bad) < 50, (mode, kf, len(bad))
            for i in bad:
                lo_ = min(int(s_log[i]), int(want_s[i]))
                assert abs(cdf[lo_] - r[i]) <= 1e-9 * cdf[-1], (i, s_log[i], want_s[i])
            del ds
    # public API, SPMD
    tc.set_distributed(True)
    c = recipes.build(tc.Circuit(n), ops)
    e = tc.templates.measurements.pauli_sum_expectation(c, pss, [w for w, _ in terms])
    want = sum(w * o.expectation_ps(ps=ps).real for w, ps in terms)
    assert abs(e - want) < 1e-4 * len(terms), (e, want)
    s = c.sample(batch=1000, allow_state=True, status=np.random.default_rng(1).random(1000), format="sample_int")
    assert s.shape == (1000,) and np.all(np.abs(ref[s]) > 0)
    psi = np.asarray(c.state())
    assert np.linalg.norm(psi - ref) / np.linalg.norm(ref) < 1e-5
    # measure on the sharded state: masks over local and global bits (basecircuit.py:359-443)
    b, pr = c.measure(0, 1, n - 1, with_prob=True, status=[0.3, 0.8, 0.5])
    oc = orc.run_gatelist(n, ops)
    # same rule on the oracle state
    p_all = np.abs(oc.state()) ** 2
    idx_all = np.arange(2**n)
    mask = value = 0
    pp = 1.0
    for k, q in enumerate((0, 1, n - 1)):
        bit = 1 << (n - 1 - q)
        pu = float(np.sum(p_all[(idx_all & (mask | bit)) == value])) / pp
        sign = 1.0 if [0.3, 0.8, 0.5][k] - pu + 0.31415926e-12 > 0 else 0.0
        assert float(b[k]) == sign, (k, b, pu)
        pp *= (1 - pu) if sign else pu
        mask |= bit
        value |= bit if sign else 0
    assert abs(float(pr) - pp) < 1e-5
    tc.set_distributed(False)
    # vmap batch sharded over the ranks (config 3's layout): one all-gather of the results
    def energy(th):
        cc = tc.Circuit(10)
        for i in range(10):
            cc.rx(i, theta=th[i])
        for i in range(9):
            cc.rzz(i, i + 1, theta=th[i] * 0.5)
        return tc.backend.real(cc.expectation_ps(z=[0, 1]) + cc.expectation_ps(x=[9]))

    th = np.random.default_rng(6).uniform(0, 6, size=(7, 10))  # 7 does not divide evenly
    got = np.asarray(tc.backend.vmap(energy)(th))
    for k in range(7):
        oo = orc.OracleCircuit(10)
        for i in range(10):
            oo.rx(i, theta=th[k, i])
        for i in range(9):
            oo.rzz(i, i + 1, theta=th[k, i] * 0.5)
        assert abs(got[k] - (oo.expectation_ps(z=[0, 1]) + oo.expectation_ps(x=[9])).real) < 1e-5, k
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
