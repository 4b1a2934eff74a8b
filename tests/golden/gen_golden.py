"""Generate the committed golden vectors under tests/golden/ from the
reference's OWN code compiled unmodified (oracle/_ref/libref.so, built by
oracle/build_ref.sh from /root/reference).  Run in the build container:

    python tests/golden/gen_golden.py

The reference tree does not travel to the GPU box; these fixtures do.
The v2 firmware vectors (v2fw_*) come from the real TIM3 ISR of mod_pdm_pwm.c +
mod_controlrate.c (oracle/ref/ref_v2_isr.c), the pwm vectors from mod_pdm.c:159-175.
The v1 carry-bit PDM (ARM inline asm, no ARM toolchain here) cannot be compiled from
the reference: its entries ("restated_v1_*") come from the oracle's restatement and
pin regressions only.  (The pwm keys keep their historical "restated_" prefix.)
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

rng = np.random.default_rng(424242)
ref = po.Ref()
orc = po.Oracle()
G = {}

# 1. linux/test_cproc.c: the reference's own generated graph, one private copy of the library
d = tempfile.mkdtemp()
priv = os.path.join(d, "libref_golden.so")
shutil.copy(po.REF_SO, priv)
r1 = po.Ref(priv)
inp = rng.integers(0, 2, 300).astype(np.uint32)
inp[100:140] = rng.integers(0, 2**32, 40, dtype=np.uint32)
msk = rng.integers(0, 4, 300).astype(np.uint32)
msk[:50] = 1
G["cproc_in"], G["cproc_mask"] = inp, msk
G["cproc_out"] = np.array([r1.test_cproc_tick(int(x), int(g))[1] for x, g in zip(inp, msk)], np.uint32)

# 2. bp5-style 3-node graph through the real acc_update / edge_update
N, F = 8, 96
gin = rng.integers(0, 3, (N, 1, F)).astype(np.uint32)
gch = rng.integers(0, 2, (N, F)).astype(np.uint32)
gst = np.zeros((N, 4), np.uint32)
G["bp5_in"], G["bp5_changed"] = gin, gch
G["bp5_out"] = ref.graph_run(po.GRAPH_BP5, 1, 2, gst, N, F, gin, gch)
G["bp5_state"] = gst

# 3. stm32f103/pdm.h pdm1..4_update
N, F = 4, 256
pin = rng.integers(0, 2**32, (N, F), dtype=np.uint32)
pdi = rng.integers(0, 1024, F).astype(np.uint32)
G["pdm_in"], G["pdm_dither"] = pin, pdi
for k in (1, 2, 3, 4):
    s = np.zeros((N, k), np.uint32)
    G["pdm%d_out" % k] = ref.pdm_run(k, s, N, F, pin, None, 24, pdi)
    G["pdm%d_state" % k] = s

# 4. v2 channel bank around the real pdm2_update: firmware config (3 channels, divider 4096)
N, F = 3, 2 * 4096 + 64
chan = np.zeros((N, 7), np.uint32)
chan[:, 0] = [2000000000, 0x40000000, 0x40000000]          # pdm_init, mod_pdm_pwm.c:147-161
prng = np.array([2463534242], np.uint32)
sp = np.array([[0x60000000, 0x80000000, 0xA0000000], [0x50000000, 0x70000000, 0xC0000000], [0x40000000] * 3], np.uint32)
chan_b, prng_b = chan.copy(), prng.copy()
duty, p1, cnt = ref.v2_isr_run(chan, int(prng[0]), 0, sp, F)      # the firmware's own ISR, once per tick
prng[0] = p1
duty_b, cnt_b = ref.pdm_v2_run(chan_b, 2, N, 3, prng_b, None, 0x3FF, 0, 12, 24, sp, F)   # the batched harness around the real pdm2_update agrees
assert np.array_equal(duty, duty_b) and np.array_equal(chan, chan_b) and np.array_equal(prng, prng_b) and cnt == cnt_b
G["v2fw_setpoints"], G["v2fw_duty_tail"], G["v2fw_state"], G["v2fw_prng"] = sp, duty[:, -256:], chan, prng
G["v2fw_duty_wsum"] = (duty.astype(np.uint64) * (np.arange(F, dtype=np.uint64) + 1)).sum(axis=1)
G["v2fw_count"] = np.array([cnt], np.uint32)
# ... and a short-divider case with 4 banks, orders 1..4
N, F = 12, 512
for k in (1, 2, 3, 4):
    chan = rng.integers(0, 2**32, (N, 5 + k), dtype=np.uint32)
    G["v2k%d_chan0" % k] = chan.copy()
    prng = np.array([1, 2, 3, 4], np.uint32)
    sp = po.pdm_setpoints(N, F // 64 + 1)
    duty, _ = ref.pdm_v2_run(chan, k, N, 3, prng, None, 0x3FF, 16, 6, 24, sp, F)
    G["v2k%d_duty" % k], G["v2k%d_state" % k], G["v2k%d_prng" % k] = duty, chan, prng

# 5. linux/synth.c: note table, synth_run / sum_tick_square
G["note_table"] = ref.note_table()
tab = G["note_table"]
v = np.zeros((2 * 64, 2), np.uint32)
v[:, 0] = np.where(rng.random(128) < 0.6, tab[rng.integers(0, 128, 128)], 0)
v[:, 1] = rng.integers(0, 2**32, 128, dtype=np.uint32)
v[64:80, 0] = 0x7FFFFFF1
G["voices0"] = v.copy()
va = v.copy(); G["saw_vec"] = ref.voice_bank_run(va, 2, 0, 64); G["saw_voices"] = va
vb = v.copy(); G["square_vec"] = ref.voice_bank_run(vb, 2, 1, 64); G["square_voices"] = vb
G["chord_vec"], G["chord_voices"] = ref.synth_play([69, 60, 127, 0], 8)

# 6. linux/synth_tools.c square_grain_proc
N, F = 8, 128
gi = rng.uniform(-1, 1, (N, F)).astype(np.float32)
th = rng.uniform(0.05, 0.5, N).astype(np.float32)
gs = rng.choice(np.array([0.0, 0.5, -0.5], np.float32), N)
G["grain_in"], G["grain_thresh"], G["grain_state0"] = gi, th, gs.copy()
G["grain_out"] = ref.square_grain_run(gs, th, N, F, gi)
G["grain_state"] = gs

# 7. v1 carry-bit PDM (restated: ARM inline asm) and pwm_update (reference lines)
N, F = 6, 256
ch = np.zeros((N, 2), np.uint32)
ch[:, 0] = [2000000000, 0x40000000, 0x60000000, 0x80000000, 0xA0000000, 0xC0000000]
prng = np.array([2463534242, 5, 9], np.uint32)
bits = orc.pdm_v1_run(ch, N, 2, prng, None, 0x0FFFFFFF, F)
G["restated_v1_bits"], G["restated_v1_state"], G["restated_v1_prng"] = np.packbits(bits, axis=1, bitorder="little"), ch, prng
ph = np.array([0, 0x123456], np.uint32); spd = np.array([256 * 13, 5000], np.uint32)
pwm_duty = np.zeros((2, 256), np.uint8)
ref._fn("pwm_run", None, [po.VP, po.VP, po.C.c_uint64, po.C.c_uint64, po.VP])(po._ptr(ph), po._ptr(spd), 2, 256, po._ptr(pwm_duty))   # mod_pdm.c:159-175
G["restated_pwm_duty"] = pwm_duty
G["restated_pwm_phase"] = ph

np.savez_compressed(os.path.join(HERE, "golden_r1.npz"), **G)
print("wrote", os.path.join(HERE, "golden_r1.npz"), os.path.getsize(os.path.join(HERE, "golden_r1.npz")), "bytes,", len(G), "arrays")
