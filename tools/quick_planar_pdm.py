"""PDM v2 at the C2 launch shape with PLANAR duty output (uint8 [ch][F]) vs TILED.  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st
N, F = 65536, 65536
ctx = st.Context(0)
d_out = ctx.dev_alloc(N * F)
rows = F // 4096
sp = np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(sp.nbytes); ctx.h2d(d_sp, sp)
for layout, name, ws, pb in ((st.TILED, "TILED", 3, 2), (st.PLANAR, "PLANAR tensor-TMA boxes", 3, 2), (st.PLANAR, "PLANAR per-lane bulk", 3, 1),
                             (st.PLANAR, "PLANAR direct stores", 3, 0), (st.TILED, "TILED", 2, 2), (st.PLANAR, "PLANAR", 2, 2)):
    for _ in (0,):
        ctx.set_option("pdm_ws", ws); ctx.set_option("pdm_planar_bulk", pb)
        b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=12, layout=layout)
        b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); ctx.sync()
        best = 1e9
        for _ in range(3):
            ctx.timer_start(); b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); best = min(best, ctx.timer_stop())
        print("%s ws=%d: %.3f ms  %.1f Gsamples/s" % (name, ws, best, N * F / best / 1e6), flush=True)
        b.free()
