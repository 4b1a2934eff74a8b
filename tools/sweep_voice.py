"""Voice bank (C4') render time by shard size, frames per launch and frames per thread.  One JSON line per variant."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st
from tools.bench_configs import note_incs

ctx = st.Context(0)
rng = np.random.default_rng(6)
for N in (4 * 1024 * 1024, 2 * 1024 * 1024, 512 * 1024):
    v = np.zeros((N, 2), np.uint32); v[:, 0] = note_incs(rng, N); v[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    b = ctx.batch(st.VOICE_BANK, N, voices_per_bus=0); b.upload_state(v)
    for F in (512, 2048):
        d_out = ctx.dev_alloc(4 * F); d_mix = ctx.dev_alloc(4 * F)
        for fpt in (0, 2, 4, 8):
            ctx.set_option("voice_fpt", fpt)
            for _ in range(3):
                b.run_dev(F, out=d_out, mix=d_mix)
            ctx.sync()
            reps = 20
            ctx.timer_start()
            for _ in range(reps):
                b.run_dev(F, out=d_out, mix=d_mix)
            ms = ctx.timer_stop() / reps
            print(json.dumps({"voices": N, "F": F, "fpt": fpt, "ms_per_launch": round(ms, 5), "us_per_512_frames": round(ms * 1e3 * 512 / F, 2),
                              "voice_samples_per_s": N * F / (ms * 1e-3)}), flush=True)
        ctx.dev_free(d_out); ctx.dev_free(d_mix)
    b.free()
ctx.set_option("voice_fpt", 0)
