"""Pass-kernel anatomy: time one pass with 1..12 two-bit blocks, for a few gathered-bit counts,
through tpass_kernel (TCB200_TMA=1) and cpass_kernel (TCB200_TMA=0).  The 1-block pass is the
data-movement floor of each kernel, the slope is its compute cost per block.
Usage: python scripts/sweep_pass.py [n] [k]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensorcircuit_b200 as tc  # noqa: E402,F401
from tensorcircuit_b200.engine import DeviceState  # noqa: E402
from tensorcircuit_b200.fusion import Block  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dtype = sys.argv[3] if len(sys.argv) > 3 else "complex64"
from tensorcircuit_b200 import _lib  # noqa: E402

T = _lib.lib.tcb200_pass_tile_bits(0 if dtype == "complex64" else 1)
rng = np.random.default_rng(0)
st = DeviceState(n, dtype)
st.init_zero()


def block(bits):
    kk = len(bits)
    u = np.linalg.qr(rng.normal(size=(2**kk, 2**kk)) + 1j * rng.normal(size=(2**kk, 2**kk)))[0]
    return Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=tuple(bits), matrix=u, batched=False, ngates=1)


for h in (0, 3, 6):
    hi = list(range(n - h, n))
    avail = list(range(T - h)) + hi
    for nops in (1, 2, 4, 6, 8, 12):
        blocks = []
        for o in range(nops):
            s = (o * k) % (len(avail) - k + 1)
            blocks.append(block(avail[s:s + k]))
        line = "h=%d nops=%2d:" % (h, nops)
        for mode in ("1", "0"):
            os.environ["TCB200_TMA"] = mode
            for _ in range(2):
                st.apply_pass_host(blocks, hi)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                st.apply_pass_host(blocks, hi)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            line += "  %s %.2f ms (%.0f GB/s)" % ("tpass" if mode == "1" else "cpass", ms, 2 * st.amp_bytes * 2**n / ms / 1e6)
        print(line, flush=True)
