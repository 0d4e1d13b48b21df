"""cProfile of K.value_and_grad on the TFIM VQE energy (python scripts/prof_grad.py [n])."""
import cProfile
import pstats
import sys

import numpy as np

sys.path.insert(0, ".")
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
layers = 4
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
vg(p)
vg(p)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    vg(p)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
