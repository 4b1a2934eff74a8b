"""Throughput of the stream processors in every layout and a few bank sizes: finds (processor, layout) pairs that sit on a slow
generic path.  One line per case.  usage: python tools/layout_matrix.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st

ctx = st.Context(0)
rng = np.random.default_rng(0)

def timed(fn, reps=3):
    fn(); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_start(); fn(); best = min(best, ctx.timer_stop())
    return best

def line(name, ms, units, bytes_per_unit):
    print("%-58s %8.3f ms  %9.3e units/s  %6.0f GB/s" % (name, ms, units / (ms * 1e-3), units * bytes_per_unit / ms / 1e6), flush=True)

# PDM v2: bank sizes x layouts (65,536 channels x 32,768 ticks)
N, F = 65536, 32768
d_out = ctx.dev_alloc(N * F)
rows = F // 4096
sp = rng.integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(sp.nbytes); ctx.h2d(d_sp, sp)
for bank in (1, 2, 3, 4, 8, 64):
    for lay in ("TILED", "PLANAR", "INTERLEAVED"):
        b = ctx.batch(st.PDM_V2, N, order=2, bank_size=bank, ctl_div_log=12, layout=getattr(st, lay))
        ms = timed(lambda: b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out)); b.free()
        line("pdm_v2 order 2 bank %d %s" % (bank, lay), ms, N * F, 1)
for order in (1, 3, 4):
    b = ctx.batch(st.PDM_V2, N, order=order, bank_size=3, ctl_div_log=12, layout=st.TILED)
    ms = timed(lambda: b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out)); b.free()
    line("pdm_v2 order %d bank 3 TILED" % order, ms, N * F, 1)
# PDM v1: bank sizes x layouts (65,536 channels x 131,072 ticks)
F1 = 131072
for bank in (1, 2, 3, 4, 8):
    for lay in ("TILED", "PLANAR", "INTERLEAVED"):
        b = ctx.batch(st.PDM_V1, N, bank_size=bank, dither_mask=0x0FFFFFFF, layout=getattr(st, lay))
        ms = timed(lambda: b.run_dev(F1, out=d_out)); b.free()
        line("pdm_v1 bank %d %s" % (bank, lay), ms, N * F1, 0.125)
ctx.dev_free(d_out); ctx.dev_free(d_sp)
# pwm (1 Mi x 1,024) and pdmK on a stream (1 Mi x 512)
N2, F2 = 1024 * 1024, 1024
d_o = ctx.dev_alloc(4 * N2 * F2)
for lay in ("TILED", "PLANAR", "INTERLEAVED"):
    b = ctx.batch(st.PWM, N2, layout=getattr(st, lay))
    b.upload_param(rng.integers(0, 2**32, (N2, 1), dtype=np.uint32))
    ms = timed(lambda: b.run_dev(F2, out=d_o)); b.free()
    line("pwm %s" % lay, ms, N2 * F2, 1)
F3 = 512
d_i = ctx.dev_alloc(4 * N2 * F3)
chunk = rng.integers(0, 2**32, (16384, F3), dtype=np.uint32)
for k in range(N2 // 16384):
    ctx.h2d(d_i + k * chunk.nbytes, chunk)
for order in (1, 2, 4):
    for lay in ("PLANAR", "INTERLEAVED"):
        b = ctx.batch(st.PDM, N2, order=order, out_shift=24, layout=getattr(st, lay))
        ms = timed(lambda: b.run_dev(F3, inp=d_i, out=d_o)); b.free()
        line("pdm%d_update on a stream %s" % (order, lay), ms, N2 * F3, 8)
for lay in ("PLANAR", "INTERLEAVED"):
    b = ctx.batch(st.ONEPOLE, N2, layout=getattr(st, lay))
    b.upload_param(rng.uniform(0.01, 0.5, (N2, 1)).astype(np.float32))
    ms = timed(lambda: b.run_dev(F3, inp=d_i, out=d_o)); b.free()
    line("onepole %s" % lay, ms, N2 * F3, 8)
    b = ctx.batch(st.WORD_CLOCK, N2, layout=getattr(st, lay))
    b.upload_param(rng.integers(10, 500, (N2, 1)).astype(np.int32).view(np.uint32))
    ms = timed(lambda: b.run_dev(F3, out=d_o)); b.free()
    line("word_clock %s" % lay, ms, N2 * F3, 4)
