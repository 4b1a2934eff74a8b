"""Warp-stall samples of one kernel by SASS opcode, and the sites that wait at a barrier, from an ncu report (needs no GPU):
    python tools/ncu_stalls_by_opcode.py gpurun_out/prof_pdm_v2_ws4.ncu-rep > profiles/r2_pdm_v2_ws4_stalls_by_opcode.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][0], rows[0][1])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError):
        return 0.0


stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(f(r, "# Samples") for r in data)
print("warp samples %d, warp instructions executed %d" % (tot, sum(f(r, "Instructions Executed") for r in data)))
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for r in data:
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    full = m.group(2) if m else src
    op = ".".join(full.split(".")[:2]) if full.startswith("IMAD.") else full.split(".")[0]
    a = agg[op]
    a[0] += f(r, "# Samples"); a[1] += f(r, "Instructions Executed")
    for c in stall_cols:
        a[2][c] += f(r, c)
print("%-12s %9s %6s %12s  reasons (share of the opcode's samples)" % ("opcode", "samples", "share", "executed"))
for op, (s, e, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:18]:
    tops = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(s, 1)) for k, v in st.most_common(4))
    print("%-12s %9.0f %5.1f%% %12.0f  %s" % (op, s, 100 * s / tot, e, tops))
print("\nsites with more than 0.1 % of the samples waiting at a barrier (the instruction after the barrier carries them):")
for n, r in enumerate(data):
    b = f(r, "stall_barrier")
    if b > tot * 1e-3:
        print("  %5.2f%%  %-44s after: %s" % (100 * b / tot, r[ix["Source"]].strip()[:44], data[n - 1][ix["Source"]].strip()[:50]))
