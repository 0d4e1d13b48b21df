"""Per-phase and per-opcode summary of an `ncu --set full --import-source on` capture of lpass_fast_kernel:
    ncu -i capture.ncu-rep --page source --csv > src.csv ;  python scripts/ncu_phase_summary.py src.csv
Phases are delimited by the two BAR.SYNC of the kernel: stage-in issue | wait for the tile | rounds | write-back.
Prints shares of the warp-stall samples and of the executed instructions, the stall reasons inside the
rounds, and the opcode mix of the rounds (profiles/r2_lpass_phase_stalls.txt is this script's output)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
seen = set()
for k, hi in enumerate(hdr_idx):
    hdr = rows[hi]
    end = hdr_idx[k + 1] - 1 if k + 1 < len(hdr_idx) else len(rows)
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[hi + 1:end] if len(r) >= len(hdr) - 5]
    sig = tuple(r[col["Instructions Executed"]] for r in body[:400])
    if sig in seen:  # the csv repeats every launch once per source view
        continue
    seen.add(sig)
    S = lambda a, b: sum(int(r[col["# Samples"]] or 0) for r in body[a:b])
    I = lambda a, b: sum(int(r[col["Instructions Executed"]] or 0) for r in body[a:b])
    bars = [i for i, r in enumerate(body) if "BAR.SYNC" in r[col["Source"]]]
    if len(bars) < 2:
        continue
    b0, b1 = bars[0], bars[-1]
    tot, ti = S(0, len(body)), I(0, len(body))
    print("== %s   %d SASS instructions, %d warp instructions executed, %d samples" % (rows[hi - 1][1][:48], len(body), ti, tot))
    print("   stage-in issue : samples %5.1f %%  instructions %5.1f %%" % (100 * S(0, b0 - 2) / tot, 100 * I(0, b0 - 2) / ti))
    print("   wait for tile  : samples %5.1f %%" % (100 * S(b0 - 2, b0 + 1) / tot))
    print("   rounds         : samples %5.1f %%  instructions %5.1f %%" % (100 * S(b0 + 1, b1 + 1) / tot, 100 * I(b0 + 1, b1 + 1) / ti))
    print("   write-back     : samples %5.1f %%  instructions %5.1f %%" % (100 * S(b1 + 1, len(body)) / tot, 100 * I(b1 + 1, len(body)) / ti))
    reg = body[b0 + 1:b1 + 1]
    names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    st = collections.Counter()
    for r in reg:
        for n in names:
            st[n] += int(r[col[n]] or 0)
    rt = sum(int(r[col["# Samples"]] or 0) for r in reg)
    print("   stalls inside the rounds: " + ", ".join("%s %.1f %%" % (n[6:], 100 * v / rt) for n, v in st.most_common(9)))
    ops, sm = collections.Counter(), collections.Counter()
    for r in reg:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]].strip())
        op = m.group(2) if m else "?"
        ops[op] += int(r[col["Instructions Executed"]] or 0)
        sm[op] += int(r[col["# Samples"]] or 0)
    it = sum(ops.values())
    print("   opcode mix of the rounds: " + ", ".join("%s %.1f %%" % (o, 100 * n / it) for o, n in ops.most_common(10)))
