"""PDM v1 at the C2 shape (65,536 channels, banks of 2, 512 Ki ticks per launch) in the three output layouts.
usage: python tools/time_pdm_v1.py [ticks] [reps] [name=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_tools_b200 as st

N = 65536
F = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = st.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
d_out = ctx.dev_alloc(N * F // 8)
for name in ("TILED", "PLANAR", "INTERLEAVED"):
    b = ctx.batch(st.PDM_V1, N, bank_size=2, dither_mask=0x0FFFFFFF, layout=getattr(st, name))
    b.run_dev(F, out=d_out); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_start(); b.run_dev(F, out=d_out); best = min(best, ctx.timer_stop())
    b.free()
    print("%-12s %.3f ms  %.3e samples/s" % (name, best, N * F / (best * 1e-3)), flush=True)
