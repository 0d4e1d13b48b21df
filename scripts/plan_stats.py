"""Host-only planning statistics of a recipe (no GPU): structure-aware fusion -> pass plan -> the
gate-pass scheduler's dry run (tcb200_gate_pass_info) per pass."""
import argparse
import os
import collections
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from tensorcircuit_b200 import _lib, fusion, gates, recipes  # noqa: E402
from tensorcircuit_b200.fusion import GateOp  # noqa: E402


def gate_ops(recipe):
    ops = []
    for name, q, p in recipe:
        g = getattr(gates, name)(**p) if p else getattr(gates, name)()
        ops.append(GateOp(q, g.matrix(), name))
    return ops


def pass_info(blocks, ids, tile_hi, nbits, dt):
    blks = [blocks[i] for i in ids]
    ks = np.asarray([len(b.bits) for b in blks], dtype=np.int32)
    bits = np.asarray([x for b in blks for x in b.bits], dtype=np.int32)
    mats = np.ascontiguousarray(np.concatenate([np.asarray(b.matrix, dtype=np.complex128).reshape(-1) for b in blks]))
    hi = np.asarray(list(tile_hi) if len(tile_hi) else [0], dtype=np.int32)
    info = np.zeros(8)
    rc = _lib.lib.tcb200_gate_pass_info(nbits, dt, len(blks), _lib.iptr(ks), _lib.iptr(bits), _lib.dptr(mats.view(np.float64)), len(tile_hi), _lib.iptr(hi), _lib.dptr(info))
    if rc:
        return None, _lib.lib.tcb200_last_error().decode()
    return info, ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=34)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--recipe", default="random")
    ap.add_argument("--max-hi", type=int, default=8)
    ap.add_argument("--max-ops", type=int, default=200)
    ap.add_argument("--nseeds", type=int, default=8)
    ap.add_argument("--kmax", type=int, default=2)
    a = ap.parse_args()
    n = a.n
    if a.recipe == "random":
        rc = recipes.random_circuit(n, a.depth, 3)
    elif a.recipe == "tfim":
        rc = recipes.tfim_vqe_circuit(n, np.random.default_rng(0).uniform(0, 1, size=(2 * a.depth, n)))
    else:
        rc = recipes.hea_circuit(n, np.random.default_rng(0).uniform(0, 1, size=(a.depth, 2, n)))
    ops = gate_ops(rc)
    t0 = time.time()
    blocks = fusion.fuse_structured(ops, n, a.kmax)
    t1 = time.time()
    print("gates", len(ops), "blocks", len(blocks), dict(collections.Counter((b.kind, len(b.bits)) for b in blocks)), "fuse %.3fs" % (t1 - t0))
    T = _lib.lib.tcb200_pass_tile_bits(0)
    cost = [0 if b.kind == "perm" else (16 if b.kind in ("diag", "mono") else 4 ** len(b.bits)) for b in blocks]
    weight = [0.0 if b.kind == "perm" else 1.0 for b in blocks]
    import os
    if os.environ.get("PLAN_W") == "fma":
        weight = [0.0 if b.kind == "perm" else (2.0 if b.kind in ("diag", "mono") else (2.0 if fusion.matrix_kind(b.matrix) == "half" else float(2 << len(b.bits)))) for b in blocks]
    if os.environ.get("PLAN_W") == "gates":
        weight = [0.0 if b.kind == "perm" else float(b.ngates) for b in blocks]
    passes = fusion.plan_passes([b.bits for b in blocks], n, T, max_hi=a.max_hi, max_ops=a.max_ops, max_mat_elems=1280, max_pass_k=3,
                                nseeds=a.nseeds, block_cost=cost, block_weight=weight,
                                block_diag=None if os.environ.get("PLAN_NODIAG") else [b.kind == "diag" for b in blocks])
    t2 = time.time()
    tot = np.zeros(8)
    for p in passes:
        info, err = pass_info(blocks, p.block_ids, p.tile_hi, n, 0)
        if info is None:
            print("pass failed:", err, len(p.block_ids))
            continue
        tot += info
        print("pass: blocks %3d hi %d rounds %2d lin %2d diag %2d dense %2d conflicts %d vec %d fma %.0f" % (
            len(p.block_ids), len(p.tile_hi), info[0], info[1], info[2], info[3], info[4], info[5], info[6]))
    print("passes", len(passes), "rounds", int(tot[0]), "conflict rounds", int(tot[4]), "vec rounds", int(tot[5]), "fma/amp", tot[6], "plan %.3fs" % (t2 - t1))


main()
