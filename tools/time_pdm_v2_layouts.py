"""PDM v2 at the C2 launch shape (65,536 channels, banks of 3, order 2) in the three duty layouts.
usage: python tools/time_pdm_v2_layouts.py [ticks per launch] [reps] [name=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st

N = 65536
F = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = st.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
d_out = ctx.dev_alloc(N * F)
rows = F // 4096
sp = np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(sp.nbytes); ctx.h2d(d_sp, sp)
for name in ("TILED", "PLANAR", "INTERLEAVED"):
    b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=12, layout=getattr(st, name))
    b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_start(); b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); best = min(best, ctx.timer_stop())
    b.free()
    print("%-12s %.3f ms  %.3e samples/s" % (name, best, N * F / (best * 1e-3)), flush=True)
