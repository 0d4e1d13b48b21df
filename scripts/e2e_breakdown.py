"""Host-side breakdown of one end-to-end step of the config-4 recipe (python scripts/e2e_breakdown.py [n])."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ops = recipes.random_circuit(n, 20, 3)
u = torch.from_numpy(np.random.default_rng(0).random(10**6)).pin_memory()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = recipes.build(tc.Circuit(n), ops)
    t1 = time.perf_counter()
    st = c._ensure_state()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    s = c.sample(batch=10**6, allow_state=True, status=u, format="sample_int")
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    del c, st
    print("n=%d  record %.1f ms | alloc+init+fuse+plan+launch (host return) %.1f ms | gpu drain %.1f ms | sample %.1f ms | total %.1f ms"
          % (n, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0)), flush=True)
