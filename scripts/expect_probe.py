"""Config-2 energy evaluation only (for ncu launch lists): python scripts/expect_probe.py [n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
params = np.random.default_rng(1).uniform(0, 2 * np.pi, [2, n])
c = recipes.build(tc.Circuit(n), recipes.tfim_vqe_circuit(n, params))
c._ensure_state()
terms = recipes.tfim_terms(n)
pss = [ps for _, ps in terms]
ws = [w for w, _ in terms]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e = tc.templates.measurements.pauli_sum_expectation(c, pss, ws)
    torch.cuda.synchronize()
    print("energy %.6f  wall %.2f ms" % (float(np.real(e)), (time.perf_counter() - t0) * 1e3), flush=True)
