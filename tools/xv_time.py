"""k_xvoice_mix render time (C4) by shard size and L2 tile size (option xvoice_vpt).  One JSON line per variant."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st
from tools.bench_configs import xvoice_records

ctx = st.Context(0)
rng = np.random.default_rng(5)
for N in (4 * 1024 * 1024, 512 * 1024):
    stt, prm = xvoice_records(rng, N)
    b = ctx.batch(st.XVOICE, N); b.upload_state(stt); b.upload_param(prm)
    F = 512
    d_mix = ctx.dev_alloc(8 * F)
    for vpt in (8, 12, 16, 64):
        ctx.set_option("xvoice_vpt", vpt)
        for _ in range(2):
            b.run_dev(F, mix=d_mix)
        ctx.sync()
        reps = 10
        ctx.timer_start()
        for _ in range(reps):
            b.run_dev(F, mix=d_mix)
        ms = ctx.timer_stop() / reps
        print(json.dumps({"xvoice_mix_voices": N, "vpt": vpt, "F": F, "ms": round(ms, 4), "voice_samples_per_s": N * F / (ms * 1e-3)}), flush=True)
    b.free(); ctx.dev_free(d_mix)
