"""Where a config-3 batch (vmap 1024 x 20 qubits, depth 4, 40 TFIM strings) spends its wall time on one GPU:
recording, flush (host return / device drain), expectation.  python scripts/config3_breakdown.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import recipes  # noqa: E402

n, B, depth = 20, 1024, 4
params = np.random.default_rng(2).uniform(0, 2 * np.pi, size=[B, depth, 2, n])
terms = recipes.tfim_terms(n)
pss, ws = [ps for _, ps in terms], [w for w, _ in terms]
marks = {}


def energy(p):
    t0 = time.perf_counter()
    c = tc.Circuit(n)
    for l in range(depth):
        for i in range(n):
            c.rx(i, theta=p[l, 0, i])
        for i in range(n - 1):
            c.rzz(i, i + 1, theta=p[l, 1, i])
        for i in range(n - 1):
            c.cnot(i, i + 1)
    t1 = time.perf_counter()
    c._ensure_state()
    t2 = time.perf_counter()
    if marks.get("sync"):
        torch.cuda.synchronize()
    t3 = time.perf_counter()
    e = tc.templates.measurements.pauli_sum_expectation(c, pss, ws)
    e = e + 0.0  # forces the lazy value
    t4 = time.perf_counter()
    marks["t"] = (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
    return e


f = tc.backend.vmap(energy)
for _ in range(2):
    f(params)
for sync in (True, False, True, False):
    marks["sync"] = sync
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f(params)
    torch.cuda.synchronize()
    w = time.perf_counter() - t0
    r, fl, dr, ex = marks["t"]
    print("sync after flush: %-5s wall %.1f ms = record %.1f + flush host %.1f + drain %.1f + expectation %.1f + rest %.1f" % (
        sync, 1e3 * w, 1e3 * r, 1e3 * fl, 1e3 * dr, 1e3 * ex, 1e3 * (w - r - fl - dr - ex)), flush=True)
