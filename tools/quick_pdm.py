"""Quick device-resident timing sweep of the PDM kernels at the C2 shape
(65,536 channels), bounded chunk.  Development tool, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st

N = 65536
F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ctx = st.Context(0)
d_out = ctx.dev_alloc(N * F)
rows = F // 4096 + 1
sp = (np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32))
d_sp = ctx.dev_alloc(sp.nbytes); ctx.h2d(d_sp, sp)

def t_run(b, reps=3, **kw):
    b.run_dev(F, **kw); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_start(); b.run_dev(F, **kw); ms = ctx.timer_stop(); best = min(best, ms)
    return best

print("v2 (pdm2 glide), N=%d F=%d" % (N, F))
for bank in (3, 4, 1):
    for tpb, persist, wps, blk in ((1, 3, 1, 64), (1, 2, 1, 64), (1, 0, 1, 64), (0, 0, 1, 64)):
            for layout in (st.TILED, st.PLANAR):
                ctx.set_option("pdm_ws", 1 if persist == 3 else 0); persist = min(persist, 2)
                ctx.set_option("pdm_tpb", tpb); ctx.set_option("pdm_block", blk); ctx.set_option("pdm_persist", persist); ctx.set_option("pdm_warps_per_smsp", wps)
                b = ctx.batch(st.PDM_V2, N, order=2, bank_size=bank, ctl_div_log=12, layout=layout)
                ms = t_run(b, ctl=d_sp, n_ctl=rows, out=d_out)
                print("bank=%d tpb=%d persist=%d wps=%d blk=%3d layout=%d : %8.3f ms  %7.3f Gsamples/s" % (bank, tpb, persist, wps, blk, layout, ms, N * F / ms / 1e6))
                b.free()
for order in (1, 3, 4):
    ctx.set_option("pdm_tpb", 1); ctx.set_option("pdm_block", 64); ctx.set_option("pdm_persist", 1); ctx.set_option("pdm_warps_per_smsp", 1); ctx.set_option("pdm_ws", 1)
    b = ctx.batch(st.PDM_V2, N, order=order, bank_size=3, ctl_div_log=12, layout=st.TILED)
    ms = t_run(b, ctl=d_sp, n_ctl=rows, out=d_out)
    print("order=%d bank=3 tpb=1 blk=64 tiled: %8.3f ms %7.3f Gsamples/s" % (order, ms, N * F / ms / 1e6))
    b.free()
print("v1 (carry-bit)")
for bank in (2, 1):
    for tpb, persist, wps, blk in ((1, 2, 1, 64), (1, 2, 2, 64), (1, 0, 1, 64), (0, 0, 1, 64)):
            for layout in (st.TILED, st.INTERLEAVED, st.PLANAR):
                ctx.set_option("pdm_tpb", tpb); ctx.set_option("pdm_block", blk); ctx.set_option("pdm_persist", persist); ctx.set_option("pdm_warps_per_smsp", wps)
                b = ctx.batch(st.PDM_V1, N, bank_size=bank, dither_mask=0x0FFFFFFF, layout=layout)
                ms = t_run(b, out=d_out)
                print("bank=%d tpb=%d persist=%d wps=%d blk=%3d layout=%d : %8.3f ms  %7.3f Gsamples/s" % (bank, tpb, persist, wps, blk, layout, ms, N * F / ms / 1e6))
                b.free()
