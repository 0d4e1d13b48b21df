"""Timing of K.value_and_grad on a TFIM VQE energy: python scripts/grad_probe.py [n] [layers]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
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
from tensorcircuit_b200 import autodiff  # noqa: E402

if len(sys.argv) > 3 and sys.argv[3] == "shift":
    autodiff._ADJOINT = False
print("adjoint sweep" if autodiff._ADJOINT else "shift rule", flush=True)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, g = vg(p)
    torch.cuda.synchronize()
    print("n=%d params=%d  value=%.6f  |grad|=%.6f  %.1f ms" % (n, p.size, float(v), float(np.linalg.norm(g)), (time.perf_counter() - t0) * 1e3), flush=True)
# spot check against a central difference of the engine's own value
h = 1e-3
e = np.zeros_like(p)
e[1, 2] = h
fd = (float(energy(p + e)) - float(energy(p - e))) / (2 * h)
print(autodiff.ADJOINT_STATS)
print("d/dp[1,2]: gradient %.6f  central difference %.6f" % (g[1, 2], fd))
