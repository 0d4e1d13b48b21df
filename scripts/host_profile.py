"""Host-side cost of a vmap batch / a gradient call without a GPU: `engine.DeviceState` is replaced by a
state whose launches are no-ops, so what is left is recording, fusion, planning and packing in Python
(the C library's own per-launch host work is not included).  Run here: python scripts/host_profile.py 3"""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import engine, recipes  # noqa: E402


class NullState(engine.DeviceState):
    def __init__(self, nbits, dtype="complex64", batch=1, device=None, buffer=None):
        self.nbits, self.dtype, self.batch = int(nbits), dtype, int(batch)
        self.dt = 0 if dtype == "complex64" else 1
        self.device = torch.device("cpu")
        self.buf = torch.zeros((1, 1), dtype=torch.complex64)
        self._ws = None

    def init_zero(self):
        pass

    def load(self, src):
        pass

    def apply_block(self, blk):
        engine.STATS["apply_launches"] += 1

    def _gate_pass_call(self, *a):
        return 0

    def _gate_pass_call_batched(self, *a):
        return 0

    def expectation_terms(self, flips, signs, nys):
        return np.zeros((self.batch, len(flips)), dtype=np.complex128)

    def row_state(self, b):
        return NullState(self.nbits, self.dtype, 1)

    def inner(self, bra, row=0, bra_row=0):
        return 0.0j

    def copy_row_from(self, row, src, src_row=0):
        pass

    def apply_pauli_sum_rows(self, *a, **k):
        pass

    def apply_csr_rows(self, *a, **k):
        pass

    def transition_local(self, bra_row, ket_row, ops):
        return np.zeros(len(ops), dtype=np.complex128)

    def norm2(self):
        return np.ones(self.batch)


def config3(B=1024, n=20, depth=4):
    params = np.random.default_rng(2).uniform(0, 2 * np.pi, size=[B, depth, 2, n])
    terms = recipes.tfim_terms(n)
    pss = [ps for _, ps in terms]
    ws = [w for w, _ in terms]

    def energy(p):
        c = tc.Circuit(n)
        for l in range(depth):
            for i in range(n):
                c.rx(i, theta=p[l, 0, i])
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=p[l, 1, i])
            for i in range(n - 1):
                c.cnot(i, i + 1)
        return tc.templates.measurements.pauli_sum_expectation(c, pss, ws)

    f = tc.backend.vmap(energy)
    return lambda: f(params)


def gradient(n=20, layers=4):
    terms = recipes.tfim_terms(n)
    pss, ws = [ps for _, ps in terms], [w for w, _ in terms]

    def energy(p):
        c = tc.Circuit(n)
        for i in range(n):
            c.h(i)
        for l in range(layers):
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=p[2 * l, i])
            for i in range(n):
                c.rx(i, theta=p[2 * l + 1, i])
        return tc.templates.measurements.pauli_sum_expectation(c, pss, ws)

    p = np.random.default_rng(0).uniform(0, 2, size=(2 * layers, n))
    vg = tc.backend.value_and_grad(energy)
    return lambda: vg(p)


if __name__ == "__main__":
    engine.DeviceState = NullState
    engine.require_cuda = lambda: None
    which = sys.argv[1] if len(sys.argv) > 1 else "3"
    run = config3() if which == "3" else gradient()
    for _ in range(3):
        run()
    t0 = time.perf_counter()
    for _ in range(10):
        run()
    print("host ms per call: %.2f" % ((time.perf_counter() - t0) * 100))
    print({k: v for k, v in engine.STATS.items() if v})
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        run()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
