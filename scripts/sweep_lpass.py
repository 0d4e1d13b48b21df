"""Anatomy of the gate pass on the GPU: ms per launch of one pass holding R rounds of two 4x4
blocks each (all 13 tile bits local, n given), against the copy-bandwidth time of the pass."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tensorcircuit_b200 import engine, fusion  # noqa: E402
from tensorcircuit_b200.fusion import Block  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(0)
    st = engine.DeviceState(n, "complex64")
    st.init_zero()
    T = 13
    hi = [n - 1 - i for i in range(8)][::-1]
    avail = list(range(5)) + hi
    out = {}
    for nrounds in (0, 1, 2, 3, 4, 6, 8, 12):
        blocks = []
        for r in range(nrounds):
            perm = rng.permutation(avail)
            for j in range(2):
                bits = tuple(sorted((int(perm[2 * j]), int(perm[2 * j + 1]))))
                u = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
                blocks.append(Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=bits, matrix=u, batched=False, ngates=1, kind="dense"))
            # serialise the rounds: a 1-bit gate on every used bit would merge; instead chain through a shared bit
        if nrounds == 0:
            bits = (0, 1)
            blocks = [Block(qubits=(n - 2, n - 1), bits=bits, matrix=np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128), batched=False, ngates=1, kind="perm")]
        before = dict(engine.STATS)
        for _ in range(2):
            st.apply_gate_pass(blocks, hi)
        torch.cuda.synchronize()
        rounds = (engine.STATS["gate_pass_rounds"] - before["gate_pass_rounds"]) // 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            st.apply_gate_pass(blocks, hi)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[str(nrounds)] = {"rounds": int(rounds), "ms": ms, "GBs": 16.0 * 2**n / ms / 1e6}
        print(nrounds, out[str(nrounds)], flush=True)
    print(json.dumps(out))


main()
