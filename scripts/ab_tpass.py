"""A/B timing of the pass kernels at one size: TCB200_TMA=1 (tpass_kernel) vs 0 (cpass_kernel).
Usage: python scripts/ab_tpass.py [n] [depth]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensorcircuit_b200 as tc  # noqa: E402
from oracle import tc_oracle as orc  # noqa: E402  (only the shared circuit recipe)
from tensorcircuit_b200 import _lib, engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ops = orc.random_circuit(n, depth, seed=0)
res = {}
for mode in ("1", "0", "1"):
    os.environ["TCB200_TMA"] = mode
    for rep in range(3):
        c = tc.Circuit(n)
        for name, q, p in ops:
            getattr(c, name)(*q, **p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0, t0 = _lib.launch_count(), _lib.lib.tcb200_tma_pass_count()
        e0.record()
        c._ensure_state()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        nl, nt = _lib.launch_count() - l0, _lib.lib.tcb200_tma_pass_count() - t0
        del c
    bytes_ = 2 * 8 * 2**n * nl
    print("TMA=%s n=%d: %.2f ms, %d launches (%d tpass), %.1f ms/launch, %.0f GB/s algorithmic" % (mode, n, ms, nl, nt, ms / nl, bytes_ / ms / 1e6), flush=True)
