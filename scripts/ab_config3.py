"""A/B of the batched (vmap) gate pass inside one process: per-element matrices through the constant bank
(lpass_fast_cbank_kernel, default) against global memory (TCB200_CBANK=0).  Config 3: vmap 1024 x 20 qubits.
Reports wall ms per batch, the host time until the call returns (launches are asynchronous) and states / s."""
import json
import os
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
f(params)
out = {}
for rep in range(3):
    for mode in ("1", "0"):
        os.environ["TCB200_CBANK"] = mode
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f(params)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out.setdefault("cbank" if mode == "1" else "global", []).append({"wall_ms": 1e3 * (t2 - t0), "host_return_ms": 1e3 * (t1 - t0)})
os.environ.pop("TCB200_CBANK", None)
for k, v in out.items():
    w = min(x["wall_ms"] for x in v)
    print(k, "wall ms", [round(x["wall_ms"], 1) for x in v], "host-return ms", [round(x["host_return_ms"], 1) for x in v], "states/s %.0f" % (B / w * 1e3))
print(json.dumps(out))

if len(sys.argv) > 1 and sys.argv[1] == "profile":
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2):
        f(params)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
