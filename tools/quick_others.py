"""Device-resident timing of the non-headline kernels at their BASELINE.json
shapes (C3a/C3b/C4/C4'/C5 of SURVEY.md 8d).  Development tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_tools_b200 as st

ctx = st.Context(0)
rng = np.random.default_rng(0)
which = sys.argv[1:] or ["voice", "grain", "gmix", "xvoice", "sweep"]
HBM = 6538.0

def timeit(fn, reps=3):
    fn(); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_start(); fn(); best = min(best, ctx.timer_stop())
    return best

NOTE = None
def note_incs(n, lo=24, hi=109):
    # equal-tempered increments, same table values as synth.c (device never recomputes them)
    tab12 = np.array([594573364, 629928536, 667386036, 707070875, 749115497, 793660223, 840853716, 890853479,
                      943826384, 999949221, 1059409296, 1122405051], np.uint32)
    notes = rng.integers(lo, hi, n)
    octave = np.where(notes < 8, 10, 9 - (notes - 8) // 12)
    idx = np.where(notes < 8, notes + 4, (notes - 8) % 12)
    return (tab12[idx] >> octave.astype(np.uint32)).astype(np.uint32)

if "voice" in which:
    N, F = 4 * 1024 * 1024, 512
    v = np.zeros((N, 2), np.uint32); v[:, 0] = note_incs(N); v[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    for G, label in ((0, "1 bus"), (64, "65536 buses of 64")):
        b = ctx.batch(st.VOICE_BANK, N, voices_per_bus=G)
        b.upload_state(v)
        nb = 1 if G == 0 else N // G
        d_out = ctx.dev_alloc(4 * nb * F); d_mix = ctx.dev_alloc(4 * nb * F)
        ms = timeit(lambda: b.run_dev(F, out=d_out, mix=d_mix))
        print("C4' voice bank %s: N=%d F=%d  %.3f ms  %.2f G voice-samples/s  (issue: %.1f%% of 37.2 T at 3 instr/voice-sample)" %
              (label, N, F, ms, N * F / ms / 1e6, 100 * 3 * N * F / (ms * 1e-3) / 37.2e12))
        ctx.dev_free(d_out); ctx.dev_free(d_mix); b.free()

if "grain" in which:
    N, F = 1024 * 1024, 256
    d_in = ctx.dev_alloc(4 * N * F); d_out = ctx.dev_alloc(4 * N * F)
    chunk = rng.uniform(-1, 1, (65536, F)).astype(np.float32)
    for k in range(N // 65536):
        ctx.h2d(d_in + k * chunk.nbytes, chunk)
    th = rng.uniform(0.05, 0.5, (N, 1)).astype(np.float32)
    for layout, label, bulk in ((st.PLANAR, "planar (register transpose)", 0), (st.PLANAR, "planar (bulk 64x3)", 1), (st.PLANAR, "planar (bulk 128x3)", 2),
                                (st.PLANAR, "planar (bulk 64x4)", 3), (st.PLANAR, "planar (bulk 32x4)", 4), (st.PLANAR, "planar (tensor TMA 64x3)", 5), (st.INTERLEAVED, "interleaved (1 grain/thread)", -1), (st.INTERLEAVED, "interleaved (4 grains/thread)", -2)):
        if bulk >= 0: ctx.set_option("grain_bulk", bulk)
        else: ctx.set_option("grain_vec4", -bulk - 1)
        b = ctx.batch(st.SQUARE_GRAIN, N, layout=layout); b.upload_param(th)
        b.run_dev(F, inp=d_in, out=d_out)       # leave the initial 0.0 state behind
        ms = timeit(lambda: b.run_dev(F, inp=d_in, out=d_out))
        print("C3a square_grain %s: N=%d F=%d  %.3f ms  %.2f G grain-samples/s  %.0f GB/s = %.1f%% of HBM (8 B/sample)" %
              (label, N, F, ms, N * F / ms / 1e6, 8 * N * F / ms / 1e6, 100 * 8 * N * F / ms / 1e6 / HBM))
        ms = timeit(lambda: b.run_dev(F, inp=d_in, out=d_in))
        print("    in place: %.3f ms  %.0f GB/s" % (ms, 8 * N * F / ms / 1e6))
        b.free()
    ctx.dev_free(d_in); ctx.dev_free(d_out)

if "gmix" in which:
    N, F = 1024 * 1024, 256
    s_rec = np.zeros((N, 2), np.uint32); s_rec[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    p_rec = np.zeros((N, 4), np.uint32)
    p_rec[:, 0] = rng.uniform(0.05, 0.5, N).astype(np.float32).view(np.uint32); p_rec[:, 1] = note_incs(N, 36, 97)
    gl = rng.integers(0, 65, N); p_rec[:, 2] = gl; p_rec[:, 3] = 64 - gl
    b = ctx.batch(st.SQUARE_GRAIN_MIX, N); b.upload_state(s_rec); b.upload_param(p_rec)
    d_out = ctx.dev_alloc(8 * F); d_mix = ctx.dev_alloc(8 * F)
    b.run_dev(F, out=d_out, mix=d_mix)      # leave the initial 0.0 state behind (steady state: every grain has flipped)
    for gen in (1, 2):
        ctx.set_option("grain_mix2", gen)
        for bps in (2, 3, 4, 6, 8):
            ctx.set_option("grain_blocks_per_sm", bps)
            ms = timeit(lambda: b.run_dev(F, out=d_out, mix=d_mix), reps=5)
            print("C3b square_grain mix gen %d (blocks/SM=%d): N=%d F=%d  %.3f ms  %.2f G grain-samples/s  (issue: %.1f%% of 37.2 T at 9 instr/grain-sample)" %
                  (gen + 1, bps, N, F, ms, N * F / ms / 1e6, 100 * 9 * N * F / (ms * 1e-3) / 37.2e12))
    ctx.set_option("grain_blocks_per_sm", 2); ctx.set_option("grain_mix2", 2)
    b.free()

def xvoice_records(N):
    prm = np.zeros((N, 8), np.uint32)
    prm[:, 0] = note_incs(N)
    prm[:, 1] = rng.uniform(0.01, 0.3, N).astype(np.float32).view(np.uint32)
    prm[:, 2] = rng.uniform(0.5, 2.0, N).astype(np.float32).view(np.uint32)
    prm[:, 3] = rng.uniform(1e-3, 1e-1, N).astype(np.float32).view(np.uint32)
    prm[:, 4] = rng.uniform(1e-3, 1e-2, N).astype(np.float32).view(np.uint32)
    prm[:, 5] = rng.integers(0, 400, N)
    g = rng.uniform(0, 1, N).astype(np.float32)
    prm[:, 6] = g.view(np.uint32); prm[:, 7] = (1 - g).astype(np.float32).view(np.uint32)
    stt = np.zeros((N, 5), np.uint32); stt[:, 0] = rng.integers(0, 2**32, N, dtype=np.uint32)
    return stt, prm

if "xvoice" in which:
    N, F = 4 * 1024 * 1024, 512
    stt, prm = xvoice_records(N)
    b = ctx.batch(st.XVOICE, N); b.upload_state(stt); b.upload_param(prm)
    d_mix = ctx.dev_alloc(8 * F)
    ms = timeit(lambda: b.run_dev(F, mix=d_mix))
    print("C4 xvoice mix: N=%d F=%d  %.3f ms  %.2f G voice-samples/s  (issue: %.1f%% of 37.2 T at 18 instr/voice-sample)" %
          (N, F, ms, N * F / ms / 1e6, 100 * 18 * N * F / (ms * 1e-3) / 37.2e12))
    b.free()

if "sweep" in which:
    N, F = 2048, 48000          # one second of the 10 s sweep
    stt, prm = xvoice_records(N)
    b = ctx.batch(st.XVOICE, N, layout=st.TILED); b.upload_state(stt); b.upload_param(prm)
    d_out = ctx.dev_alloc(8 * N * F)
    ms = timeit(lambda: b.run_dev(F, out=d_out), reps=2)
    print("C5 sweep raw (thread per variant): N=%d F=%d  %.3f ms  %.3f G variant-frames/s  %.1f GB/s = %.2f%% of HBM" %
          (N, F, ms, N * F / ms / 1e6, 8 * N * F / ms / 1e6, 100 * 8 * N * F / ms / 1e6 / HBM))
    b.free()
    ctx.dev_free(d_out)
    # the full C5 shard: 2,048 variants x 480,000 frames (10 s at 48 kHz), time-parallel scan path
    F = 480000
    d_out = ctx.dev_alloc(8 * N * F)
    for layout, lname in ((st.TILED, "TILED"), (st.PLANAR, "PLANAR")):
        for chunk, groups in ((0, 1), (0, 4), (0, 8), (128, 8), (512, 8), (1024, 8)):
            ctx.set_option("xvoice_chunk", chunk); ctx.set_option("xvoice_groups", groups)
            b = ctx.batch(st.XVOICE, N, layout=layout, mode=st.XVOICE_SCAN); b.upload_state(stt); b.upload_param(prm)
            ms = timeit(lambda: b.run_dev(F, out=d_out), reps=2)
            print("C5 sweep raw (time-parallel scan, %s, chunk=%d, groups=%d): N=%d F=%d  %.3f ms  %.3f G variant-frames/s  %.1f GB/s = %.2f%% of HBM" %
                  (lname, chunk, groups, N, F, ms, N * F / ms / 1e6, 8 * N * F / ms / 1e6, 100 * 8 * N * F / ms / 1e6 / HBM))
            b.free()
    ctx.set_option("xvoice_chunk", 0); ctx.set_option("xvoice_groups", 0)
