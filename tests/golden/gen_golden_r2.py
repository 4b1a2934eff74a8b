"""Round-2 golden vectors (tests/golden/golden_r2.npz), like gen_golden.py from the reference's OWN
code compiled unmodified (oracle/_ref/libref.so, oracle/build_ref.sh).  Run in the build container:

    python tests/golden/gen_golden_r2.py

ext_*: the extension processors of include/cproc_ext.h -- DEF_PROC bodies compiled against the reference's
generic/cproc.h (oracle/ref/ref_cproc.c: ref_graph_run_ext) on node tables that mix them with acc / edge, and the
graph texts tests/golden/ext_voice.cproc / ext_chain.cproc compiled as C with the reference's PROC macros
(oracle/ref/ref_ext_voice.c).  The reference has no such processors: these pin the C definition, not reference
behaviour.
pixi_*: the PIXI demo LFO bank, stm32f103/pixi.c:279,282-285 (`inc = adc[0] >> 5; dac = (dac + inc) & 0xFFF`,
12 DAC channels), for several knob positions and start values.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

rng = np.random.default_rng(20261018)
ref = po.Ref()
G = {}

adc = np.array([0, 31, 32, 1000, 4095, 65535, 40000], np.uint16)        # 12-bit ADC readings and beyond (uint16 >> 5)
dac0 = rng.integers(0, 0x1000, (len(adc), 12)).astype(np.uint16)
dac0[0] = 0
dac0[1] = 0xFFF
T = 300
trace = np.zeros((len(adc), T, 12), np.uint16)
dac1 = dac0.copy()
for k, a in enumerate(adc):
    trace[k] = ref.pixi_lfo_run(dac1[k], int(a), T)
G["pixi_adc0"], G["pixi_dac0"], G["pixi_trace"], G["pixi_dac1"] = adc, dac0, trace, dac1

# extension processors: three node tables through the DEF_PROC bodies
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
from test_graph_ext import random_ext_graph, random_params, state_words  # noqa: E402
for k in range(3):
    N, F, n_in = 6, 160, 2
    rows = random_ext_graph(rng, [5, 9, 14][k], n_in, masks=k == 2)
    prm = random_params(rng, rows, N)
    inp = rng.integers(0, 2**32, (N, n_in, F), dtype=np.uint32)
    inp[:, 1] = rng.uniform(-1, 1, (N, F)).astype(np.float32).view(np.uint32)
    chg = rng.integers(0, 8, (N, F)).astype(np.uint32) if k == 2 else np.zeros((0,), np.uint32)
    st = np.zeros((N, state_words(rows)), np.uint32)
    outs = list(range(len(rows)))[-3:]
    out = ref.graph_run_ext(rows, n_in, outs, st, prm, N, F, inp, chg if k == 2 else None)
    G["ext%d_rows" % k] = np.array([[r[0], r[1], r[2], r[3] if len(r) > 3 else 0] for r in rows], np.int64)
    G["ext%d_param" % k] = prm if prm is not None else np.zeros((N, 0), np.uint32)
    G["ext%d_in" % k], G["ext%d_changed" % k], G["ext%d_out" % k], G["ext%d_state" % k] = inp, chg, out, st
# the graph texts compiled as C (one instance each, from zero state; private copies: node state is function-static)
import shutil, tempfile  # noqa: E402
d = tempfile.mkdtemp()
priv = os.path.join(d, "libref_golden_r2.so")
shutil.copy(po.REF_SO, priv)
r1 = po.Ref(priv)
F = 600
mod = rng.integers(0, 1 << 22, (1, F)).astype(np.uint32)
G["extvoice_in"], G["extvoice_out"] = mod, r1.ext_text_run(0, mod, None, F, 2)
cin = np.zeros((2, F), np.uint32)
cin[0] = rng.integers(0, 2, F); cin[1] = rng.integers(0, 1 << 20, F)
cch = rng.integers(0, 4, F).astype(np.uint32)
G["extchain_in"], G["extchain_changed"], G["extchain_out"] = cin, cch, r1.ext_text_run(1, cin, cch, F, 3)
gin = np.repeat(rng.uniform(0, 1, F // 64 + 1).astype(np.float32), 64)[:F].view(np.uint32).reshape(1, F).copy()
G["extgain_in"], G["extgain_out"] = gin, r1.ext_text_run(2, gin, None, F, 1)

np.savez_compressed(os.path.join(HERE, "golden_r2.npz"), **G)
print("wrote golden_r2.npz,", len(G), "arrays")
