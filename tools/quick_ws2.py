"""Development sweep of the PDM v2 warp-specialised kernels at the C2 launch shape
(65,536 channels x F ticks, banks of 3, TILED).  Every variant is first checked
bit for bit (duty bytes, channel state, PRNG state) against the CPU oracle on a
small shape and against the first-generation kernel on the full shape.
Development tool, not the bench."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st
from oracle import pyoracle as po

N = 65536
F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
variants = [(1, 0, 1, 2)] + [(2, f, c, ns) for f in (0, 1, 2) for c in (1, 2, 4) for ns in (2, 4)]
ctx = st.Context(0)


def setopt(ws, form, chains, slots):
    ctx.set_option("pdm_slots", slots)
    ctx.set_option("pdm_ws", ws)
    ctx.set_option("pdm_form", form)
    ctx.set_option("pdm_chains", chains)


# ---- parity on a small shape against the oracle
orc = po.Oracle()
n, f = 3 * 700 + 2, 4096 * 3
chan0 = np.zeros((n, 7), np.uint32)
chan0[:, 5:7] = np.random.default_rng(1).integers(0, 2**32, (n, 2), dtype=np.uint32)
nb = (n + 2) // 3
prng0 = (np.arange(nb) * 2654435761 + 12345).astype(np.uint32) | 1
sp = po.pdm_setpoints(n, f // 4096)
ca, pa = chan0.copy(), prng0.copy()
want, _ = orc.pdm_v2_run(ca, 2, n, 3, pa, None, 0x3FF, 0, 12, 24, sp, f)
for ws, form, chains, slots in variants:
    setopt(ws, form, chains, slots)
    b = ctx.batch(st.PDM_V2, n, order=2, bank_size=3, ctl_div_log=12, layout=st.TILED)
    b.upload_state(chan0); b.upload_bank(prng0, 0)
    out = np.zeros(n * f, np.uint8)
    b.run(f, ctl=sp, out=out)
    got = out.reshape(f // 16, n, 16).transpose(1, 0, 2).reshape(n, f)
    ok = np.array_equal(got, want) and np.array_equal(b.download_state(), ca) and np.array_equal(b.download_bank()[0], pa)
    print("parity ws=%d form=%d chains=%d slots=%d : %s" % (ws, form, chains, slots, "bit-exact" if ok else "MISMATCH"), flush=True)
    b.free()

# ---- timing + cross-check at the C2 launch shape
d_out = ctx.dev_alloc(N * F)
rows = F // 4096 + 1
spf = np.random.default_rng(0).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
d_sp = ctx.dev_alloc(spf.nbytes); ctx.h2d(d_sp, spf)
ref_crc = None
host = np.zeros(N * min(F, 4096), np.uint8)
for ws, form, chains, slots in variants:
    setopt(ws, form, chains, slots)
    b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=12, layout=st.TILED)
    b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); ctx.sync()
    ctx.d2h(host, d_out + (N * F - host.nbytes))            # the last ticks of the launch
    crc = (zlib.crc32(host.tobytes()), zlib.crc32(b.download_state().tobytes()), zlib.crc32(b.download_bank()[0].tobytes()))
    if ref_crc is None:
        ref_crc = crc
    best = 1e9
    for _ in range(4):
        ctx.timer_start(); b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out); best = min(best, ctx.timer_stop())
    print("ws=%d form=%d chains=%d slots=%d : %8.3f ms  %8.1f Gsamples/s  %s" % (ws, form, chains, slots, best, N * F / best / 1e6,
          "same as ws1" if crc == ref_crc else "DIFFERS from ws1"), flush=True)
    b.free()
