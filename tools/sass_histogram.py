"""Per-kernel SASS opcode histograms of the built library (needs no GPU):
    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt
For every kernel of synth_tools_b200/build/*.o: instruction count, the async / TMA / barrier opcodes that prove what the kernel
is made of (UTMALDG / UTMASTG tensor TMA, UBLKCP bulk copies, SYNCS mbarriers, BAR named barriers, FFMA2 / FADD2 / FMUL2 packed
fp32, no HMMA / UTCMMA: nothing here is a contraction) and the ten most frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "synth_tools_b200", "build")
KEY = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "BAR", "FFMA2", "FADD2", "FMUL2", "LDGSTS", "ATOM", "RED", "HMMA", "UTCMMA", "SHFL", "MEMBAR", "CCTL"]

tot = collections.Counter()
for o in sorted(os.listdir(OBJ)):
    if not o.endswith(".o"):
        continue
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, o)], capture_output=True, text=True).stdout
    fn = None
    per = collections.OrderedDict()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            per[fn][m.group(1)] += 1
    print("== %s" % o)
    for fn, c in per.items():
        n = sum(c.values())
        keys = "  ".join("%s %d" % (k, sum(v for op, v in c.items() if op.startswith(k))) for k in KEY if any(op.startswith(k) for op in c))
        top = " ".join("%s:%d" % kv for kv in c.most_common(10))
        print("%-70s %6d instr | %s | %s" % (fn[:70], n, keys or "-", top))
        for k in KEY:
            tot[k] += sum(v for op, v in c.items() if op.startswith(k))
print("== library totals: " + "  ".join("%s %d" % (k, tot[k]) for k in KEY))
