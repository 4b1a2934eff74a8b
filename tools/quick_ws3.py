"""PDM v2 dynamic-schedule kernel (k_pdm_v2_ws3) at the C2 launch shape: persistent blocks
per SM x slice length, cross-checked (CRC of the last ticks, channel state, PRNG state)
against the static ws2 launch.  Development tool, not the bench."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st

N = 65536
F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ctx = st.Context(0)
d_out = ctx.dev_alloc(N * F)
rows = F // 4096 + 1
spf = np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(spf.nbytes); ctx.h2d(d_sp, spf)
host = np.zeros(N * min(F, 4096), np.uint8)
ref_crc = None
variants = [(2, 4, 64, 1, 2, 0)] + [(3, 4, sb, form, ch, fma) for fma in (0, 1) for form, ch in ((1, 2), (2, 2), (1, 4), (1, 1)) for sb in (32, 64, 128)]
for ws, ctas, sb, form, chains, fma in variants:
    ctx.set_option("pdm_prng_fma", fma)
    ctx.set_option("pdm_ws", ws); ctx.set_option("pdm_ctas_per_sm", ctas); ctx.set_option("pdm_slice_batches", sb)
    ctx.set_option("pdm_form", form); ctx.set_option("pdm_chains", chains)
    b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=12, layout=st.TILED)
    b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); ctx.sync()
    ctx.d2h(host, d_out + (N * F - host.nbytes))
    crc = (zlib.crc32(host.tobytes()), zlib.crc32(b.download_state().tobytes()), zlib.crc32(b.download_bank()[0].tobytes()))
    if ref_crc is None:
        ref_crc = crc
    best = 1e9
    for _ in range(5):
        ctx.timer_start(); b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); best = min(best, ctx.timer_stop())
    print("ws=%d ctas/SM=%d slice=%d batches form=%d chains=%d prng_fma=%d : %8.3f ms  %8.1f Gsamples/s  %s" %
          (ws, ctas, sb, form, chains, fma, best, N * F / best / 1e6, "same as ws2" if crc == ref_crc else "DIFFERS from ws2"), flush=True)
    b.free()
