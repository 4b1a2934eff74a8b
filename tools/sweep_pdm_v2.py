"""k_pdm_v2_ws4 at the C2 launch shape (65,536 channels, banks of 3): batch length x persistent blocks per SM x slice length.
Every variant writes the same bytes (CRC of the slab against the first).  One JSON line per variant.
usage: python tools/sweep_pdm_v2.py [ticks per launch] [reps] [planar]"""
import json
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st

N = 65536
F = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
layout = st.PLANAR if len(sys.argv) > 3 and sys.argv[3] == "planar" else st.TILED
ctx = st.Context(0)
d_out = ctx.dev_alloc(N * F)
rows = F // 4096
sp = np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(sp.nbytes); ctx.h2d(d_sp, sp)
host = np.zeros(N * F, np.uint8) if F <= 65536 else None
ref = None
for tlog, ctas, sb in [(7, 4, 64), (6, 4, 64), (7, 5, 64), (7, 3, 64), (7, 4, 32), (7, 4, 128), (7, 4, 256)]:
    for k, v in (("pdm_tlog", tlog), ("pdm_ctas_per_sm", ctas), ("pdm_slice_batches", sb)):
        ctx.set_option(k, v)
    b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=12, layout=layout)
    b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); ctx.sync()
    crc = None
    if host is not None:
        ctx.d2h(host, d_out)
        crc = zlib.crc32(host)
        ref = crc if ref is None else ref
    best = 1e9
    for _ in range(reps):
        ctx.timer_start(); b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); best = min(best, ctx.timer_stop())
    b.free()
    print(json.dumps({"tlog": tlog, "ctas_per_sm": ctas, "slice_batches": sb, "ms": round(best, 4),
                      "samples_per_s": N * F / (best * 1e-3), "same_bytes": None if crc is None else crc == ref}), flush=True)
