"""SURVEY 5 row 2 (sanitizers), CPU side: the oracle's C restatement built with -fsanitize=undefined (gcc's UBSan: shifts,
alignment, out-of-bounds indexing of fixed arrays, float -> int conversions, null dereferences; signed wrap-around is defined
by the oracle's -fwrapv contract) and driven through every entry point on randomised inputs in a subprocess that aborts on
the first report.  (compute-sanitizer is closed on the GPU pool: profiles/r2_sanitizer_note.txt.)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
from oracle import pyoracle as po
o = po.Oracle.__new__(po.Oracle); po._Lib.__init__(o, %(so)r)
rng = np.random.default_rng(2)
# graphs: every node kind, masks, several outputs
rows = [(po.NODE_EDGE, -1, 1), (po.NODE_ACC, 0, 3), (po.node_glide(5), 1, 0xFFFFFFFF), (po.node_pdm(3, 24), 2, 0xFFFFFFFF, -2),
        (po.NODE_PHASOR_F, -1, 2), (po.NODE_SVF, 4, 0xFFFFFFFF), (po.NODE_ENV, 5, 1), (po.NODE_ONEPOLE, 6, 0xFFFFFFFF), (po.NODE_GAIN, 1, 0xFFFFFFFF),
        (po.NODE_ASFLOAT, -2, 0xFFFFFFFF), (po.NODE_PHASOR_F, po.SRC_ZERO, 0xFFFFFFFF)]
N, F = 7, 300
sw = sum(po.node_words(r[0]) for r in rows); pw = sum(po.node_param_words(r[0]) for r in rows)
st = rng.integers(0, 2**32, (N, sw), dtype=np.uint32); prm = rng.integers(0, 2**32, (N, pw), dtype=np.uint32)
inp = rng.integers(0, 2**32, (N, 2, F), dtype=np.uint32); chg = rng.integers(0, 4, (N, F)).astype(np.uint32)
o.graph_run_ext(rows, 2, [3, 7, 8, 10], st, prm, N, F, inp, chg)
o.graph_run(rows[:4], 2, 3, st[:, :12].copy(), N, F, inp, chg)
o.graph_run_multi(rows[:4], 2, [0, 3], st[:, :12].copy(), N, F, inp)
# pdm family, all orders and extreme shifts
for k in (1, 2, 3, 4):
    for sh in (0, 24, 31):
        o.pdm_run(k, rng.integers(0, 2**32, (N, k), dtype=np.uint32), N, F, inp[:, 0].copy(), None, sh, rng.integers(0, 2**32, F, dtype=np.uint32))
    for ctl in (1, 12, 24):
        chan = rng.integers(0, 2**32, (N, 5 + k), dtype=np.uint32)
        o.pdm_v2_run(chan, k, N, 3, rng.integers(1, 2**32, 3, dtype=np.uint32), None, 0x3FF, 0, ctl, 24, po.pdm_setpoints(N, F), F)
o.pdm_v1_run(rng.integers(0, 2**32, (N, 2), dtype=np.uint32), N, 2, rng.integers(1, 2**32, 4, dtype=np.uint32), None, 0x0FFFFFFF, 256)
o.pwm_run(rng.integers(0, 2**24, N).astype(np.uint32), rng.integers(0, 2**24, N).astype(np.uint32), N, F)
# voices, grains, extension voice, one-pole, word clock
v = rng.integers(0, 2**32, (128, 2), dtype=np.uint32)
for mode in (0, 1):
    o.voice_bank_run(v.copy(), 128, 64, mode, F)
o.square_grain_run(np.zeros(N, np.float32), rng.uniform(0, 1, N).astype(np.float32), N, F, rng.uniform(-1, 1, (N, F)).astype(np.float32))
o.square_grain_mix_run(np.zeros(N, np.float32), rng.uniform(0, 1, N).astype(np.float32), rng.integers(0, 2**32, N, dtype=np.uint32),
                       rng.integers(0, 2**32, N, dtype=np.uint32), rng.integers(0, 65, N).astype(np.uint8), rng.integers(0, 65, N).astype(np.uint8), N, F)
xs = np.zeros(N, po.xvoice_state_dtype); xs["phase"] = rng.integers(0, 2**32, N, dtype=np.uint32); xs["t"] = 0xFFFFFFF0
xp = np.zeros(N, po.xvoice_param_dtype); xp["inc"] = rng.integers(0, 2**32, N, dtype=np.uint32); xp["f"] = 0.2; xp["q"] = 1; xp["env_attack"] = 0.1; xp["env_release"] = 0.01; xp["gate_frames"] = 100; xp["gl"] = 1
o.xvoice_run(xs, xp, N, F)
o.onepole_run(np.zeros(N, np.float32), rng.uniform(0, 1, N).astype(np.float32), N, F, rng.uniform(-1, 1, (N, F)).astype(np.float32))
o.word_clock_run(np.zeros((N, 2), np.int32), rng.integers(1, 50, N).astype(np.int32), N, F)
[o.note_to_inc(n) for n in range(128)]
print("ubsan driver ok")
'''


def test_oracle_under_ubsan(tmp_path):
    so = str(tmp_path / "liboracle_ubsan.so")
    subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-g", "-fwrapv", "-ffp-contract=off", "-mfma", "-fopenmp", "-fPIC", "-shared", "-Wall",
                           "-fsanitize=undefined", "-fno-sanitize-recover=all", os.path.join(ROOT, "oracle", "cproc_oracle.c"), "-o", so, "-lm"])
    env = dict(os.environ, UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    res = subprocess.run([sys.executable, "-c", DRIVER % {"root": ROOT, "so": so}], capture_output=True, text=True, env=env)
    assert res.returncode == 0 and "ubsan driver ok" in res.stdout and "runtime error" not in res.stderr, res.stderr[-3000:]
