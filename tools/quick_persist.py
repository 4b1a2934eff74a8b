"""Development: thread-per-bank kernels (persistent / plain) vs the warp-specialised one at the C2 launch shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st
N, F = 65536, 65536
ctx = st.Context(0)
d_out = ctx.dev_alloc(N * F)
rows = F // 4096 + 1
spf = np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(spf.nbytes); ctx.h2d(d_sp, spf)
for name, opts in (("ws2", dict(pdm_ws=2)), ("persist wps1", dict(pdm_ws=0, pdm_persist=2, pdm_warps_per_smsp=1)),
                   ("persist wps2", dict(pdm_ws=0, pdm_persist=2, pdm_warps_per_smsp=2)),
                   ("plain tpb blk32", dict(pdm_ws=0, pdm_persist=0, pdm_block=32)), ("plain tpb blk64", dict(pdm_ws=0, pdm_persist=0, pdm_block=64)),
                   ("plain tpb blk128", dict(pdm_ws=0, pdm_persist=0, pdm_block=128))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=12, layout=st.TILED)
    b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); ctx.sync()
    best = 1e9
    for _ in range(3):
        ctx.timer_start(); b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); best = min(best, ctx.timer_stop())
    print("%-18s %8.3f ms  %8.1f Gsamples/s" % (name, best, N * F / best / 1e6), flush=True)
    b.free()
