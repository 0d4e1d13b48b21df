"""A few launches of one synthetic pass (for ncu): python scripts/one_pass.py n nops h k [dtype]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensorcircuit_b200 as tc  # noqa: E402,F401
from tensorcircuit_b200.engine import DeviceState  # noqa: E402
from tensorcircuit_b200.fusion import Block  # noqa: E402

n, nops, h, k = (int(x) for x in sys.argv[1:5])
dtype = sys.argv[5] if len(sys.argv) > 5 else "complex64"
from tensorcircuit_b200 import _lib  # noqa: E402

T = _lib.lib.tcb200_pass_tile_bits(0 if dtype == "complex64" else 1)
rng = np.random.default_rng(0)
st = DeviceState(n, dtype)
st.init_zero()
hi = list(range(n - h, n))
avail = list(range(T - h)) + hi
blocks = []
for o in range(nops):
    s = (o * k) % (len(avail) - k + 1)
    bits = avail[s:s + k]
    u = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))[0]
    blocks.append(Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=tuple(bits), matrix=u, batched=False, ngates=1))
for _ in range(3):
    st.apply_pass_host(blocks, hi)
torch.cuda.synchronize()
