"""One gate pass holding K pairs of 4x4 blocks at n qubits, launched a few times (for ncu)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tensorcircuit_b200 import engine  # noqa: E402
from tensorcircuit_b200.fusion import Block  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rng = np.random.default_rng(0)
st = engine.DeviceState(n, "complex64")
st.init_zero()
hi = [n - 1 - i for i in range(8)][::-1]
avail = list(range(5)) + hi
blocks = []
for r in range(k):
    perm = rng.permutation(avail)
    for j in range(2):
        bits = tuple(sorted((int(perm[2 * j]), int(perm[2 * j + 1]))))
        u = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
        blocks.append(Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=bits, matrix=u, batched=False, ngates=1, kind="dense"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.apply_gate_pass(blocks, hi)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    st.apply_gate_pass(blocks, hi)
e1.record()
torch.cuda.synchronize()
print("n", n, "k", k, "rounds", engine.STATS["gate_pass_rounds"] // (reps + 1), "ms/launch", e0.elapsed_time(e1) / reps)
