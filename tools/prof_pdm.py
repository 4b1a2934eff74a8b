"""Three launches of the headline kernel at the C2 launch shape (65,536 channels x
65,536 ticks, banks of 3, TILED out) for `ncu --set full`.  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st

N, F = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
which = sys.argv[2] if len(sys.argv) > 2 else "v2"
ctx = st.Context(0)
for kv in sys.argv[3:]:                      # name=value tuning options
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
d_out = ctx.dev_alloc(N * F)
rows = F // 4096
sp = np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(sp.nbytes); ctx.h2d(d_sp, sp)
if which == "v2":
    b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=12, layout=st.TILED)
    for _ in range(3):
        b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out)
else:
    b = ctx.batch(st.PDM_V1, N, bank_size=2, dither_mask=0x0FFFFFFF, layout=st.TILED)
    for _ in range(3):
        b.run_dev(F, out=d_out)
ctx.sync()
print("ok", ctx.launches)
