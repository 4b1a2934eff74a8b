"""Device-resident timing of the non-headline BASELINE.json configurations at their full
shapes (C3a, C3b, C4, C4', C5), each against the roofline that bounds it.
Used by bench.py (key "other_configs", N=1 only) and by tools/quick_others.py.

Timing: CUDA events on the context stream (cproc_cuda_timer_*), best of `reps` after one
warm-up run; every buffer is larger than L2 or the kernel is issue bound (stated per row).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ISSUE_PEAK = 128.0 * 148 * 1.965e9       # thread-instr/s at sm_max_mhz (tools/ubench_int.cu: 128 /clk/SM)
TAB12 = np.array([594573364, 629928536, 667386036, 707070875, 749115497, 793660223, 840853716, 890853479,
                  943826384, 999949221, 1059409296, 1122405051], np.uint32)     # linux/synth.c:78-95 (gcc, verified)


def note_incs(rng, n, lo=24, hi=109):
    notes = rng.integers(lo, hi, n)
    octave = np.where(notes < 8, 10, 9 - (notes - 8) // 12)
    idx = np.where(notes < 8, notes + 4, (notes - 8) % 12)
    return (TAB12[idx] >> octave.astype(np.uint32)).astype(np.uint32)


def xvoice_records(rng, N):
    prm = np.zeros((N, 8), np.uint32)
    prm[:, 0] = note_incs(rng, N)
    prm[:, 1] = rng.uniform(0.01, 0.3, N).astype(np.float32).view(np.uint32)
    prm[:, 2] = rng.uniform(0.5, 2.0, N).astype(np.float32).view(np.uint32)
    prm[:, 3] = rng.uniform(1e-3, 1e-1, N).astype(np.float32).view(np.uint32)
    prm[:, 4] = rng.uniform(1e-3, 1e-2, N).astype(np.float32).view(np.uint32)
    prm[:, 5] = rng.integers(0, 400, N)
    g = rng.uniform(0, 1, N).astype(np.float32)
    prm[:, 6] = g.view(np.uint32); prm[:, 7] = (1 - g).astype(np.float32).view(np.uint32)
    stt = np.zeros((N, 5), np.uint32); stt[:, 0] = rng.integers(0, 2**32, N, dtype=np.uint32)
    return stt, prm


def _time(ctx, fn, reps, before=None):
    """best of `reps` CUDA-event timings of fn; `before` (untimed) runs ahead of every call, e.g. to put the state back"""
    if before:
        before()
    fn(); ctx.sync()
    best = 1e9
    for _ in range(reps):
        if before:
            before(); ctx.sync()
        ctx.timer_start(); fn(); best = min(best, ctx.timer_stop())
    return best


def _hbm(name, units, unit_name, ms, bytes_per_unit, hbm_peak, note):
    gbs = bytes_per_unit * units / ms / 1e6
    return {"config": name, "value": units / (ms * 1e-3), "unit": unit_name + "/s", "ms": ms, "bound": "hbm",
            "achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak, "algorithmic_bytes_per_unit": bytes_per_unit,
            "note": note}


def _issue(name, units, unit_name, ms, instr_per_unit, note):
    rate = instr_per_unit * units / (ms * 1e-3)
    return {"config": name, "value": units / (ms * 1e-3), "unit": unit_name + "/s", "ms": ms, "bound": "issue",
            "achieved_tinstr_s": rate / 1e12, "peak_tinstr_s": ISSUE_PEAK / 1e12, "frac": rate / ISSUE_PEAK,
            "algorithmic_instr_per_unit": instr_per_unit, "note": note}


# --------------------------------------------------------------------------- CPU baselines
# The reference's own loops (oracle/_ref, kind "reference": compiled here from /root/reference by oracle/build_ref.sh) or,
# where the processor is an extension or the reference code is ARM assembly, the oracle port (kind "port"), timed on this
# box's host cores on a bounded sample of the row's workload (about 1 s of CPU work each).  Reported beside the GPU
# number; not the optimisation target.
def _cpu_time(fn, units, min_s=0.6):
    import time
    fn()                                           # warm: page in, first touch
    reps, dt = 0, 0.0
    t0 = time.perf_counter()
    while dt < min_s and reps < 50:
        fn(); reps += 1
        dt = time.perf_counter() - t0
    return units * reps / dt


def cpu_baselines():
    """name -> {"value", "unit", "cores", "kind", "sample"} for every row of run_all."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    po.set_threads(cores)
    orc = po.Oracle()
    try:
        ref = po.Ref()
    except Exception:
        ref = None
    rng = np.random.default_rng(11)
    out = {}

    def put(name, value, unit, used, kind, sample):
        out[name] = {"value": value, "unit": unit, "cores": used, "kind": kind, "sample": sample}

    # C1 / bp5 / C1 long: the generated graphs (generic/cproc.h acc_update / edge_update behind ref_graph_run: serial)
    lib, kind = (ref, "reference") if ref else (orc, "port")
    for name, rows in (("c1", po.GRAPH_TEST_CPROC), ("graph_bp5", po.GRAPH_BP5)):
        N, F = 4096, 256
        inp = rng.integers(0, 2, (N, 1, F), dtype=np.uint32)
        stt = np.zeros((N, sum(po.node_words(r[0]) for r in rows)), np.uint32)
        v = _cpu_time(lambda: lib.graph_run(rows, 1, len(rows) - 1, stt, N, F, inp), N * F)
        put(name, v, "ticks/s", 1 if kind == "reference" else cores, kind, "%d instances x %d ticks (cproc.h acc_update / edge_update, one thread)" % (N, F))
    out["c1_long"] = out["c1"]
    # C2 v1: carry-bit PDM (ARM assembly in the reference: oracle port, OpenMP over banks)
    N, F = 2048 * max(1, cores // 8), 16384
    ch = np.zeros((N, 2), np.uint32); ch[:, 0] = rng.integers(0x40000000, 0xC0000000, N, dtype=np.uint32)
    pr = (np.arange(N // 2) + 1).astype(np.uint32)
    v = _cpu_time(lambda: orc.pdm_v1_run(ch, N, 2, pr, None, 0x0FFFFFFF, F), N * F)
    put("c2_v1", v, "samples/s", cores, "port", "%d channels x %d ticks, banks of 2 (restated adds / rrx of mod_pdm.c:214-244), OpenMP" % (N, F))
    # C3a: square_grain_proc itself
    N, F = 4096 * max(1, cores // 8), 256
    inp = rng.uniform(-1, 1, (N, F)).astype(np.float32); th = rng.uniform(0.05, 0.5, N).astype(np.float32)
    stt = np.zeros(N, np.float32); o = np.zeros((N, F), np.float32)
    v = _cpu_time(lambda: lib.square_grain_run(stt, th, N, F, inp, out=o), N * F)
    put("c3a", v, "grain-samples/s", cores, kind, "%d grains x %d frames (square_grain_proc, synth_tools.c:85-100), OpenMP over grains" % (N, F))
    # C3b: phasor -> trigger -> integer mix (oracle port, one thread: the mix is a serial sum)
    N, F = 16384, 256
    s0 = np.zeros(N, np.float32); ph = rng.integers(0, 2**32, N, dtype=np.uint32); inc = note_incs(rng, N, 36, 97)
    gl = rng.integers(0, 65, N).astype(np.uint8); gr = (64 - gl).astype(np.uint8)
    v = _cpu_time(lambda: orc.square_grain_mix_run(s0, th[:1].repeat(N), ph, inc, gl, gr, N, F), N * F)
    put("c3b", v, "grain-samples/s", 1, "port", "%d grains x %d frames, one thread" % (N, F))
    # C4: extension voice (oracle port, one thread)
    N, F = 4096, 512
    sx = np.zeros(N, po.xvoice_state_dtype); px = np.zeros(N, po.xvoice_param_dtype)
    px["inc"] = note_incs(rng, N); px["f"] = 0.1; px["q"] = 1.0; px["env_attack"] = 0.01; px["env_release"] = 0.001; px["gate_frames"] = 200
    px["gl"] = 0.5; px["gr"] = 0.5
    v = _cpu_time(lambda: orc.xvoice_run(sx, px, N, F, want_raw=False), N * F)
    put("c4", v, "voice-samples/s", 1, "port", "%d voices x %d frames, stereo mix, one thread" % (N, F))
    vr = _cpu_time(lambda: orc.xvoice_run(sx, px, N, F, want_mix=False), N * F)
    put("c5", vr, "variant-frames/s", 1, "port", "%d variants x %d frames, raw stereo out, one thread (sequential in time)" % (N, F))
    # the same voice as a graph of DEF_PROC processors (include/cproc_ext.h compiled against the reference's cproc.h), one thread
    N, F = 2048, 512
    grows = [(po.NODE_PHASOR_F, po.SRC_ZERO, 0xFFFFFFFF), (po.NODE_SVF, 0, 0xFFFFFFFF), (po.NODE_ENV, 1, 0xFFFFFFFF), (po.NODE_GAIN, 2, 0xFFFFFFFF), (po.NODE_GAIN, 2, 0xFFFFFFFF)]
    gp = np.ascontiguousarray(px[:N].view(np.uint32).reshape(N, 8)); gs = np.zeros((N, 9), np.uint32)
    v = _cpu_time(lambda: lib.graph_run_ext(grows, 0, [3, 4], gs, gp, N, F), N * F)
    put("graph_voice", v, "ticks/s", 1 if kind == "reference" else cores, "port", "%d instances x %d ticks (cproc_ext.h DEF_PROC bodies behind a node table%s)" % (N, F, ", one thread" if kind == "reference" else ", OpenMP"))
    # C4': synth_run / sum_tick_saw of linux/synth.c:169-202, 64-voice synths
    NS, F = 2048 * max(1, cores // 8), 512
    vv = np.zeros((NS * 64, 2), np.uint32); vv[:, 0] = note_incs(rng, NS * 64); vv[:, 1] = rng.integers(0, 2**32, NS * 64, dtype=np.uint32)
    if ref:
        v = _cpu_time(lambda: ref.voice_bank_run(vv, NS, 0, F), NS * 64 * F)
        put("c4p", v, "voice-samples/s", cores, "reference", "%d synths x 64 voices x %d frames (synth_run), OpenMP over synths" % (NS, F))
    else:
        v = _cpu_time(lambda: orc.voice_bank_run(vv, NS * 64, 64, 0, F), NS * 64 * F)
        put("c4p", v, "voice-samples/s", cores, "port", "%d synths x 64 voices x %d frames, OpenMP over synths" % (NS, F))
    # pdm2_update on an input stream
    N, F = 2048 * max(1, cores // 8), 4096
    inp = rng.integers(0, 2**32, (N, F), dtype=np.uint32); stt = np.zeros((N, 2), np.uint32)
    v = _cpu_time(lambda: lib.pdm_run(2, stt, N, F, inp, None, 24, None), N * F)
    put("pdm_raw", v, "samples/s", cores, kind, "%d channels x %d samples (pdm2_update, pdm.h:32-40), OpenMP over channels" % (N, F))
    return out


def c1(st, ctx):
    """BASELINE.json configs[0]: the linux/test_cproc.c chain (edge -> acc), 1 voice, 64-frame blocks,
    750 blocks (1 s at 48 kHz) through the host-buffer call a JACK period makes: latency, not throughput."""
    import time
    rng = np.random.default_rng(1)
    F, blocks = 64, 750
    b = ctx.batch(st.GRAPH, 1, nodes=[(st.NODE_EDGE, -1, 1), (st.NODE_ACC, 0, 1)])
    inp = rng.integers(0, 2, (blocks, 1, 1, F), dtype=np.uint32)
    out = np.zeros((1, F), np.uint32)
    for k in range(20):
        b.run(F, inp=inp[k], out=out)
    t0 = time.perf_counter()
    for k in range(blocks):
        b.run(F, inp=inp[k], out=out)
    us = (time.perf_counter() - t0) / blocks * 1e6
    b.free()
    return {"config": "C1 test_cproc chain (edge -> acc), 1 voice x 64-frame blocks, 750 blocks through cproc_cuda_run (host buffers)",
            "value": us, "unit": "us/block", "bound": "launch latency", "block_period_us": 64 / 48000 * 1e6,
            "note": "per block: memcpy 256 B to pinned staging, one cudaGraphLaunch (the generated kernel reads and writes the staging), stream sync, memcpy back; a 64-frame period at 48 kHz is 1333 us"}


def c2_v1(st, ctx, reps=3):
    """SURVEY 8d C2 variant v1: the carry-bit PDM of mod_pdm.c:198-264, banks of 2 sharing dither, 1 bit per sample."""
    N, F = 65536, 512 * 1024
    d_out = ctx.dev_alloc(N * F // 8)
    b = ctx.batch(st.PDM_V1, N, bank_size=2, dither_mask=0x0FFFFFFF, layout=st.TILED)
    sp = np.random.default_rng(2).integers(0x40000000, 0xC0000000, (N, 2), dtype=np.uint32)
    sp[:, 1] = 0
    b.upload_state(sp)
    ms = _time(ctx, lambda: b.run_dev(F, out=d_out), reps)
    b.free(); ctx.dev_free(d_out)
    return _issue("C2 v1 carry-bit PDM (mod_pdm.c), 65,536 ch x 512 Ki ticks, banks of 2, packed bits out (4 GiB)", N * F, "samples", ms, 7.0,
                  "SURVEY 8d: 4 int instr per channel-sample + 6 PRNG instr per bank-tick / 2 channels; 1/8 B per sample out")


def c2_layout(st, ctx, layout, reps=3):
    """C2 PDM v2 at the headline's channel count with the duty bytes in another layout (the headline is TILED): PLANAR rows through
    tensor-TMA boxes, INTERLEAVED [tick][ch] -- the order the ISR of mod_pdm_pwm.c produces them -- through block tiles."""
    N, F, L = 65536, 32768, 12
    d_out = ctx.dev_alloc(N * F)
    rows = F >> L
    sp = np.random.default_rng(6).integers(0x40000000, 0xC0000000, (rows, N), dtype=np.uint32)
    d_sp = ctx.dev_alloc(sp.nbytes); ctx.h2d(d_sp, sp)
    b = ctx.batch(st.PDM_V2, N, order=2, bank_size=3, ctl_div_log=L, layout=getattr(st, layout.upper()))
    ms = _time(ctx, lambda: b.run_dev(F, ctl=d_sp, n_ctl=rows, out=d_out), reps)
    b.free(); ctx.dev_free(d_out); ctx.dev_free(d_sp)
    return _issue("C2 PDM v2, 65,536 ch x 32,768 ticks per launch, %s duty out" % ("PLANAR [ch][F]" if layout == "planar" else "INTERLEAVED [tick][ch]"),
                  N * F, "samples", ms, 10.0, "the headline's modulator in the other two duty layouts (DESIGN 4.1)")


def c3a(st, ctx, hbm_peak, reps=5, layout="planar"):
    rng = np.random.default_rng(3)
    N, F = 1024 * 1024, 256
    d_in = ctx.dev_alloc(4 * N * F); d_out = ctx.dev_alloc(4 * N * F)
    chunk = rng.uniform(-1, 1, (65536, F)).astype(np.float32)
    for k in range(N // 65536):
        ctx.h2d(d_in + k * chunk.nbytes, chunk)
    b = ctx.batch(st.SQUARE_GRAIN, N, layout=st.PLANAR if layout == "planar" else st.INTERLEAVED)
    b.upload_param(rng.uniform(0.05, 0.5, (N, 1)).astype(np.float32))
    ms = _time(ctx, lambda: b.run_dev(F, inp=d_in, out=d_out), reps)
    b.free(); ctx.dev_free(d_in); ctx.dev_free(d_out)
    return _hbm("C3a square_grain raw, 1 Mi grains x 256 frames, %s float in/out" % layout, N * F, "grain-samples", ms, 8.0, hbm_peak,
                "4 B in + 4 B out per grain-sample; 1 GiB in + 1 GiB out, larger than L2")


def c3b(st, ctx, reps=5):
    rng = np.random.default_rng(4)
    N, F = 1024 * 1024, 256
    s_rec = np.zeros((N, 2), np.uint32); s_rec[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    p_rec = np.zeros((N, 4), np.uint32)
    p_rec[:, 0] = rng.uniform(0.05, 0.5, N).astype(np.float32).view(np.uint32); p_rec[:, 1] = note_incs(rng, N, 36, 97)
    gl = rng.integers(0, 65, N); p_rec[:, 2] = gl; p_rec[:, 3] = 64 - gl
    b = ctx.batch(st.SQUARE_GRAIN_MIX, N); b.upload_state(s_rec); b.upload_param(p_rec)
    d_out = ctx.dev_alloc(8 * F); d_mix = ctx.dev_alloc(8 * F)
    b.run_dev(F, out=d_out, mix=d_mix)          # leave the initial 0.0 state behind (steady state)
    ms = _time(ctx, lambda: b.run_dev(F, out=d_out, mix=d_mix), reps)
    b.free(); ctx.dev_free(d_out); ctx.dev_free(d_mix)
    return _issue("C3b square_grain phasor -> trigger -> stereo integer mix, 1 Mi grains x 256 frames", N * F, "grain-samples", ms, 9.0,
                  "SURVEY 8d: 9 algorithmic instr per grain-sample; timed region = memset + mix kernel + int->float kernel")


def c4(st, ctx, reps=3):
    rng = np.random.default_rng(5)
    N, F = 4 * 1024 * 1024, 512
    stt, prm = xvoice_records(rng, N)
    b = ctx.batch(st.XVOICE, N); b.upload_state(stt); b.upload_param(prm)
    d_mix = ctx.dev_alloc(8 * F)
    # the block that follows note-on: every voice crosses its gate (attack -> release) inside the 512 frames; the state goes back
    # to t = 0 before every timed launch (untimed).  Left alone, the second launch on finds every voice released: `steady`.
    ms = _time(ctx, lambda: b.run_dev(F, mix=d_mix), reps, before=lambda: b.upload_state(stt))
    ms_steady = _time(ctx, lambda: b.run_dev(F, mix=d_mix), reps)
    b.free(); ctx.dev_free(d_mix)
    row = _issue("C4 poly voice (phasor + SVF + AR envelope + pan), 4 Mi voices x 512 frames, float stereo mix", N * F, "voice-samples", ms, 18.0,
                 "SURVEY 8d: 18 algorithmic instr per voice-sample; deterministic fixed-order mix; timed on the block after note-on (gates at random "
                 "frames 0..399 of the block: per-tick attack / release selection); `steady` = the following blocks, every voice released")
    row["steady"] = {"ms": ms_steady, "value": N * F / (ms_steady * 1e-3)}
    return row


def c4p(st, ctx, reps=5):
    rng = np.random.default_rng(6)
    N, F = 4 * 1024 * 1024, 512
    v = np.zeros((N, 2), np.uint32); v[:, 0] = note_incs(rng, N); v[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    b = ctx.batch(st.VOICE_BANK, N, voices_per_bus=0); b.upload_state(v)
    d_out = ctx.dev_alloc(4 * F); d_mix = ctx.dev_alloc(4 * F)
    ms = _time(ctx, lambda: b.run_dev(F, out=d_out, mix=d_mix), reps)
    b.free(); ctx.dev_free(d_out); ctx.dev_free(d_mix)
    return _issue("C4' reference voice bank (sum_tick_saw, linux/synth.c:169-181), 4 Mi voices x 512 frames, int32 mix", N * F, "voice-samples", ms, 3.0,
                  "3 instr per voice-sample in the closed-form kernel (IMAD state + t*inc, SHF >> 4, IADD; SURVEY 8d a-7 counts 4 for the "
                  "reference's read-modify-write loop); bit-exact at any reduction order")


def c5(st, ctx, hbm_peak, reps=3, layout="tiled"):
    rng = np.random.default_rng(7)
    N, F = 2048, 480000
    stt, prm = xvoice_records(rng, N)
    d_out = ctx.dev_alloc(8 * N * F)
    b = ctx.batch(st.XVOICE, N, layout=st.TILED if layout == "tiled" else st.PLANAR, mode=st.XVOICE_SCAN)
    b.upload_state(stt); b.upload_param(prm)
    ms = _time(ctx, lambda: b.run_dev(F, out=d_out), reps)
    b.free(); ctx.dev_free(d_out)
    return _hbm("C5 patch sweep shard, 2,048 variants x 480,000 frames (10 s @ 48 kHz) stereo float raw out, %s, time-parallel scan" % layout,
                N * F, "variant-frames", ms, 8.0, hbm_peak, "8 B out per variant-frame (7.86 GB per launch, larger than L2); "
                "timed region = envelope walk + closed-form zero-state pass + scan + render")


def scaling_rows(st, ctx, torch, stream, dev, rank, world, dist=None):
    """The BASELINE.json configurations that shard over the GPUs of a box, at N = `world` (1 included, so that the per-N
    lines of a scaling run carry their own 1-GPU denominator).  STRONG scaling: the total work is fixed (4 Mi voices,
    1 Mi grains, 16,384 variants) and split in contiguous shards.  C4', C4, C3b end in the shared mix bus (the only
    exchange of the path); C5 has none.  Device time (CUDA events on the render stream), max over ranks.  Every mix is
    checked on rank 0 against the single-device oracle (integer buses bit for bit, the float bus within the stated
    tolerance and identical on all ranks)."""
    from synth_tools_b200 import shard
    from oracle import pyoracle as po
    orc = po.Oracle()
    po.set_threads(os.cpu_count() or 1)
    rows = []

    def timed(fn, reps, after=None):
        for _ in range(3):
            fn()
        if after:
            after()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        if after:
            after()
        e1.record(stream)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(flag):
        t = torch.tensor([int(bool(flag))], device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    bus = shard.connect_bus(st.Bus(ctx, 8192, world, rank)) if world > 1 else None

    # ---- C4': reference voice bank, int32 mix bus ------------------------------------------------------------------
    try:
        N, F = 4 * 1024 * 1024, 512
        rng = np.random.default_rng(99)
        v = np.zeros((N, 2), np.uint32)
        v[:, 0] = note_incs(rng, N, 0, 128); v[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
        lo, hi = shard.shard_range(N, rank, world)
        b = ctx.batch(st.VOICE_BANK, hi - lo, voices_per_bus=0)

        def oracle_blocks(n_frames):                      # 64 pseudo-buses in parallel, integer wrap-around sum of the rows: exact
            w = v.copy()
            isum, _ = orc.voice_bank_run(w, N, N // 64, 0, n_frames, want_vec=False)
            tot = isum.view(np.uint32).sum(axis=0, dtype=np.uint64).astype(np.uint32)
            return tot.view(np.int32), (tot.view(np.int32).astype(np.float32) * np.float32(2.0 ** -32))

        want_i, want_f = oracle_blocks(4 * F) if rank == 0 else (None, None)
        res, exact = {}, {}
        imix = [torch.zeros(4 * F, dtype=torch.int32, device=dev) for _ in range(2)]
        out = [torch.zeros(4 * F, dtype=torch.float32, device=dev) for _ in range(2)]
        k = [0]

        def check(name, frames, blocks):
            """render `blocks` launches of `frames` from the initial state and compare the concatenation with the oracle"""
            b.upload_state(np.ascontiguousarray(v[lo:hi]))
            gi = [torch.zeros(frames, dtype=torch.int32, device=dev) for _ in range(blocks)]
            go = [torch.zeros(frames, dtype=torch.float32, device=dev) for _ in range(blocks)]
            for q in range(blocks):
                if name == "nccl":
                    b.run_dev(frames, mix=gi[q].data_ptr())
                    if dist is not None:
                        dist.all_reduce(gi[q])
                    b.mix_to_float(gi[q].data_ptr(), go[q].data_ptr(), frames)
                else:
                    b.run_dev(frames, mix=gi[q].data_ptr(), out=go[q].data_ptr())
            if bus is not None:
                bus.flush()
            torch.cuda.synchronize()
            ok = True
            if rank == 0:
                ci = torch.cat(gi).cpu().numpy(); co = torch.cat(go).cpu().numpy()
                ok = np.array_equal(ci, want_i[:frames * blocks]) and np.array_equal(co.view(np.uint32), want_f[:frames * blocks].view(np.uint32))
            return all_ok(ok and (bus is None or bus.status() == 0))

        def block(frames):
            def fn():
                s_ = k[0] & 1
                k[0] += 1
                b.run_dev(frames, mix=imix[s_].data_ptr(), out=out[s_].data_ptr())
            return fn

        def block_nccl():
            b.run_dev(F, mix=imix[0].data_ptr())
            dist.all_reduce(imix[0][:F])
            b.mix_to_float(imix[0].data_ptr(), out[0].data_ptr(), F)

        variants = [("one 512-frame block per launch, exchange inside the launch", 1, F, 4), ("one 512-frame block per launch, exchange pipelined beside the next launch", 2, F, 4),
                    ("four 512-frame blocks per launch, exchange pipelined", 2, 4 * F, 1)]
        for name, mode, frames, blocks in variants:
            if bus is not None:
                bus.attach(b, mode)
            exact[name] = check(name, frames, blocks)
            res[name] = timed(block(frames), 40, after=(bus.flush if bus is not None else None)) * F / frames
        if bus is not None:
            bus.detach(b)
            exact["NCCL all-reduce + conversion kernel (baseline)"] = check("nccl", F, 2)
            res["NCCL all-reduce + conversion kernel (baseline)"] = timed(block_nccl, 40)
        b.free()
        rows.append({"config": "C4' reference voice bank (sum_tick_saw, linux/synth.c:169-181), 4 Mi voices x 512-frame blocks over %d GPU%s, int32 mix bus" % (world, "s" if world > 1 else ""),
                     "n_gpus": world, "scaling": "strong", "unit": "voice-samples/s", "ms_per_512_frame_block": res,
                     "value": {kk: N * F / (vv * 1e-3) for kk, vv in res.items()}, "bit_exact": exact,
                     "note": "the bus of ALL ranks (int32 words and the float scale) compared with the single-device oracle on rank 0; "
                             "1 GPU: the same launches without a bus"})
    except Exception as e:
        rows.append({"config": "C4'", "error": "%s: %s" % (type(e).__name__, e)})

    # ---- C4: extension voices, float stereo bus (rank-order float sum) ---------------------------------------------
    try:
        N, F, NC = 4 * 1024 * 1024, 512, 64 * 1024 + 5
        rng = np.random.default_rng(5)
        stt, prm = xvoice_records(rng, N)
        # parity on a 64 Ki-voice subset against the oracle (float mix: tolerance; identical bits on all ranks)
        lo, hi = shard.shard_range(NC, rank, world)
        sa = np.ascontiguousarray(stt[:NC]).view(po.xvoice_state_dtype).reshape(NC).copy()
        pa = np.ascontiguousarray(prm[:NC]).view(po.xvoice_param_dtype).reshape(NC)
        _, want = orc.xvoice_run(sa, pa, NC, F, want_raw=False) if rank == 0 else (None, None)
        b = ctx.batch(st.XVOICE, hi - lo)
        b.upload_state(np.ascontiguousarray(stt[lo:hi])); b.upload_param(np.ascontiguousarray(prm[lo:hi]))
        mix = torch.zeros(2 * F, dtype=torch.float32, device=dev)
        if bus is not None:
            bus.attach(b, 1)
        b.run_dev(F, mix=mix.data_ptr())
        torch.cuda.synchronize()
        ok, same = True, True
        if rank == 0:
            g64, w64 = mix.cpu().numpy().astype(np.float64), np.asarray(want, np.float64).reshape(-1)
            ok = bool(np.abs(g64 - w64).max() <= 1e-5 * np.abs(w64).max() and
                      10 * np.log10((w64 ** 2).sum() / max(((g64 - w64) ** 2).sum(), 1e-300)) >= 120.0)
        if dist is not None:
            allmix = [torch.empty_like(mix) for _ in range(world)]
            dist.all_gather(allmix, mix)
            same = all(torch.equal(allmix[0], m) for m in allmix)
        # four blocks in ONE launch ([2][4 F] mix, one exchange) == four launches of one block, bit for bit (offline renders)
        lo_c, hi_c = lo, hi
        b.upload_state(np.ascontiguousarray(stt[lo_c:hi_c]))
        four = [torch.zeros(2 * F, dtype=torch.float32, device=dev) for _ in range(4)]
        for q in range(4):
            b.run_dev(F, mix=four[q].data_ptr())
        if bus is not None:
            bus.flush()
        torch.cuda.synchronize()
        b.upload_state(np.ascontiguousarray(stt[lo_c:hi_c]))
        one4 = torch.zeros(2 * 4 * F, dtype=torch.float32, device=dev)
        b.run_dev(4 * F, mix=one4.data_ptr())
        if bus is not None:
            bus.flush()
        torch.cuda.synchronize()
        same4 = all(torch.equal(one4.view(2, 4, F)[:, q, :].reshape(-1).view(torch.int32), four[q].view(torch.int32)) for q in range(4))
        if bus is not None:
            bus.detach(b)
        b.free()
        lo, hi = shard.shard_range(N, rank, world)
        b = ctx.batch(st.XVOICE, hi - lo)
        b.upload_state(np.ascontiguousarray(stt[lo:hi])); b.upload_param(np.ascontiguousarray(prm[lo:hi]))
        mixes = [torch.zeros(2 * F, dtype=torch.float32, device=dev) for _ in range(2)]
        mixes4 = [torch.zeros(2 * 4 * F, dtype=torch.float32, device=dev) for _ in range(2)]
        k = [0]

        def xblock4():
            s_ = k[0] & 1
            k[0] += 1
            b.run_dev(4 * F, mix=mixes4[s_].data_ptr())

        def xblock():
            s_ = k[0] & 1
            k[0] += 1
            b.run_dev(F, mix=mixes[s_].data_ptr())

        def xnccl():
            b.run_dev(F, mix=mixes[0].data_ptr())
            dist.all_reduce(mixes[0])

        res = {}
        for name, mode in (("exchange inside the launch", 1), ("exchange pipelined beside the next launch", 2)):
            if bus is not None:
                bus.attach(b, mode)
            res[name] = timed(xblock, 10, after=(bus.flush if bus is not None else None))
        res["four 512-frame blocks per launch, exchange pipelined"] = timed(xblock4, 5, after=(bus.flush if bus is not None else None)) / 4
        if bus is not None:
            bus.detach(b)
            res["NCCL all-reduce (baseline)"] = timed(xnccl, 10)
        b.free()
        rows.append({"config": "C4 poly voice (phasor + SVF + AR envelope + pan), 4 Mi voices x 512-frame blocks over %d GPU%s, float stereo mix bus" % (world, "s" if world > 1 else ""),
                     "n_gpus": world, "scaling": "strong", "unit": "voice-samples/s", "ms_per_512_frame_block": res,
                     "value": {kk: N * F / (vv * 1e-3) for kk, vv in res.items()},
                     "within_tolerance": all_ok(ok), "identical_bits_on_all_ranks": all_ok(same), "four_blocks_per_launch_bit_identical_to_four_launches": all_ok(same4),
                     "note": "parity on a %d-voice subset: <= 1e-5 of peak and >= 120 dB SNR against the C oracle; the bus is a float sum in rank order" % NC})
    except Exception as e:
        rows.append({"config": "C4", "error": "%s: %s" % (type(e).__name__, e)})

    # ---- C3b: square_grain mix, integer bus in units of 2^-7 --------------------------------------------------------
    try:
        N, F = 1024 * 1024, 256
        rng = np.random.default_rng(4)
        s_rec = np.zeros((N, 2), np.uint32); s_rec[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
        p_rec = np.zeros((N, 4), np.uint32)
        th = rng.uniform(0.05, 0.5, N).astype(np.float32)
        p_rec[:, 0] = th.view(np.uint32); p_rec[:, 1] = note_incs(rng, N, 36, 97)
        gl = rng.integers(0, 65, N); p_rec[:, 2] = gl; p_rec[:, 3] = 64 - gl
        lo, hi = shard.shard_range(N, rank, world)
        b = ctx.batch(st.SQUARE_GRAIN_MIX, hi - lo)
        b.upload_state(np.ascontiguousarray(s_rec[lo:hi])); b.upload_param(np.ascontiguousarray(p_rec[lo:hi]))
        imix = [torch.zeros(2 * F, dtype=torch.int32, device=dev) for _ in range(2)]
        out = [torch.zeros(2 * F, dtype=torch.float32, device=dev) for _ in range(2)]
        # block 0 against the oracle
        b.run_dev(F, out=out[0].data_ptr(), mix=imix[0].data_ptr())
        if bus is not None:
            bus.allreduce(imix[0].data_ptr(), 2 * F, out_dev=out[0].data_ptr(), scale=st.Bus.SCALE_GRAIN)
        torch.cuda.synchronize()
        ok = True
        if rank == 0:
            wi, wf = orc.square_grain_mix_run(np.zeros(N, np.float32), th, s_rec[:, 1].copy(), p_rec[:, 1].copy(), gl.astype(np.uint8), (64 - gl).astype(np.uint8), N, F)
            ok = np.array_equal(imix[0].cpu().numpy().reshape(2, F), wi) and np.array_equal(out[0].cpu().numpy().reshape(2, F).view(np.uint32), wf.view(np.uint32))
        k = [0]

        def gblock():
            s_ = k[0] & 1
            k[0] += 1
            if bus is not None:
                bus.wait(s_)
            b.run_dev(F, out=out[s_].data_ptr(), mix=imix[s_].data_ptr())
            if bus is not None:
                bus.begin(s_, imix[s_].data_ptr(), 2 * F, out_dev=out[s_].data_ptr(), scale=st.Bus.SCALE_GRAIN)

        def gafter():
            if bus is not None:
                bus.wait(0); bus.wait(1)

        ms = timed(gblock, 20, after=gafter)
        b.free()
        rows.append({"config": "C3b square_grain phasor -> trigger -> stereo integer mix, 1 Mi grains x 256-frame blocks over %d GPU%s" % (world, "s" if world > 1 else ""),
                     "n_gpus": world, "scaling": "strong", "unit": "grain-samples/s", "ms_per_block": ms, "value": N * F / (ms * 1e-3),
                     "bit_exact": all_ok(ok and (bus is None or bus.status() == 0)),
                     "note": "exchange: the bus kernel over NVLink peer memory on its own stream, overlapped with the next block's render"})
    except Exception as e:
        rows.append({"config": "C3b", "error": "%s: %s" % (type(e).__name__, e)})
    if bus is not None:
        bus.destroy()

    # ---- C5: patch sweep, 16,384 variants x 480,000 frames raw stereo out, no collective -------------------------------
    try:
        NV, F, SH = 16384, 480000, 2048
        lo, hi = shard.shard_range(NV, rank, world)
        rng = np.random.default_rng(7)
        stt, prm = xvoice_records(rng, NV)
        d_out = ctx.dev_alloc(8 * SH * F)
        batches = []
        for a in range(lo, hi, SH):
            e = min(hi, a + SH)
            bb = ctx.batch(st.XVOICE, e - a, layout=st.TILED, mode=st.XVOICE_SCAN)
            bb.upload_state(np.ascontiguousarray(stt[a:e])); bb.upload_param(np.ascontiguousarray(prm[a:e]))
            batches.append(bb)

        def sweep():
            for bb in batches:
                bb.run_dev(F, out=d_out)

        ms = timed(sweep, 2)
        for bb in batches:
            bb.free()
        ctx.dev_free(d_out)
        rows.append({"config": "C5 patch sweep, 16,384 variants x 480,000 frames (10 s @ 48 kHz) stereo float raw out over %d GPU%s, TILED, time-parallel scan" % (world, "s" if world > 1 else ""),
                     "n_gpus": world, "scaling": "strong", "unit": "variant-frames/s", "ms": ms, "value": NV * F / (ms * 1e-3),
                     "variants_per_gpu": hi - lo, "launch_shards_per_gpu": len(batches), "collective": None,
                     "achieved_gbs_per_gpu": 8.0 * (hi - lo) * F / ms / 1e6,
                     "note": "independent variants: no exchange; each GPU renders its variants in shards of 2,048 (7.86 GB of output each, the slab is reused); parity: tests/test_gpu_fullsize.py"})
    except Exception as e:
        rows.append({"config": "C5", "error": "%s: %s" % (type(e).__name__, e)})
    return rows


def graph_bp5(st, ctx, hbm_peak, reps=3, layout="planar"):
    """the reference's generated bp5 graph (edge -> acc -> acc, stm32f103/bp5_plugin.c:1-9),
    parsed from its text and rendered by the kernel compiled for it."""
    rng = np.random.default_rng(8)
    rows, n_in, out_node, _ = st.graph_parse("""#define CPROC_NB_INPUTS 1
        void cproc_update(w *input, w g) {
            PROC_COND(g&0b1, n1, edge, NULL, NULL, .in = input[0]);
            PROC_COND(g&0b1, n2, acc, NULL, NULL, .in = n1.out);
            PROC_COND(g&0b1, n3, acc, NULL, NULL, .in = n2.out);
            cproc_output(3, n3.out); }""")
    N, F = 4 * 1024 * 1024, 256
    d_in = ctx.dev_alloc(4 * N * F); d_out = ctx.dev_alloc(4 * N * F)
    chunk = rng.integers(0, 2, (65536, F), dtype=np.uint32)
    for k in range(N // 65536):
        ctx.h2d(d_in + k * chunk.nbytes, chunk)
    b = ctx.batch(st.GRAPH, N, nodes=rows, n_inputs=n_in, out_node=out_node, layout=st.PLANAR if layout == "planar" else st.INTERLEAVED)
    ms = _time(ctx, lambda: b.run_dev(F, inp=d_in, out=d_out), reps)
    b.free(); ctx.dev_free(d_in); ctx.dev_free(d_out)
    return _hbm("generated cproc graph bp5 (edge -> acc -> acc), 4 Mi instances x 256 ticks, %s uint32 in/out, NVRTC kernel" % layout, N * F, "ticks", ms, 8.0,
                hbm_peak, "4 B in + 4 B out per tick; 4 GiB in + 4 GiB out")


VOICE_GRAPH_TEXT = """#define CPROC_NB_INPUTS 0
    void cproc_update(w *input, w g) {
        PROC(osc, phasor_f, NULL, &osc_p);
        PROC(flt, svf, NULL, &flt_p, .in = osc.out);
        PROC(amp, env, NULL, &amp_p, .in = flt.out);
        PROC(left, gain, NULL, &left_p, .in = amp.out);
        PROC(right, gain, NULL, &right_p, .in = amp.out);
        cproc_output_f(0, left.out);
        cproc_output_f(1, right.out); }"""


def graph_voice(st, ctx, hbm_peak, reps=3, layout="planar"):
    """SURVEY 8 a-X as cproc processors: the C4 voice (phasor_f -> svf -> env -> pan gains, include/cproc_ext.h) written as
    generated graph text, parsed and rendered by the kernel NVRTC compiles for it; raw stereo float out, no input stream."""
    rng = np.random.default_rng(12)
    g = st.graph_parse_ex(VOICE_GRAPH_TEXT)
    N, F = 2 * 1024 * 1024, 512
    stt, prm = xvoice_records(rng, N)
    gst = np.zeros((N, 9), np.uint32); gst[:, 1] = stt[:, 0]
    d_out = ctx.dev_alloc(8 * N * F)
    b = ctx.batch(st.GRAPH, N, nodes=g["rows"], n_inputs=0, out_node=g["out_nodes"], layout=st.PLANAR if layout == "planar" else st.INTERLEAVED)
    b.upload_state(gst); b.upload_param(prm)
    ms = _time(ctx, lambda: b.run_dev(F, out=d_out), reps)
    b.free(); ctx.dev_free(d_out)
    return _hbm("C4 voice as a generated cproc graph (phasor_f -> svf -> env -> 2 x gain, cproc_ext.h), 2 Mi instances x 512 ticks, %s float out x 2, NVRTC kernel" % layout,
                N * F, "ticks", ms, 8.0, hbm_peak, "0 B in + 8 B out per tick (8 GiB out); bit-identical to k_xvoice raw output (tests/test_graph_ext.py)")


def pdm_raw(st, ctx, hbm_peak, reps=3, layout="planar"):
    """pdm2_update (stm32f103/pdm.h) on a caller-supplied input stream, uint32 in / out, planar [ch][F] or interleaved [F][ch]
    (random words: the layout only decides who reads which)."""
    rng = np.random.default_rng(10)
    N, F = 1024 * 1024, 1024
    d_in = ctx.dev_alloc(4 * N * F); d_out = ctx.dev_alloc(4 * N * F)
    chunk = rng.integers(0, 2**32, (16384, F), dtype=np.uint32)
    for k in range(N // 16384):
        ctx.h2d(d_in + k * chunk.nbytes, chunk)
    b = ctx.batch(st.PDM, N, order=2, out_shift=24, layout=st.PLANAR if layout == "planar" else st.INTERLEAVED)
    ms = _time(ctx, lambda: b.run_dev(F, inp=d_in, out=d_out), reps)
    b.free(); ctx.dev_free(d_in); ctx.dev_free(d_out)
    how = "planar uint32 in/out (tensor-TMA staging)" if layout == "planar" else "interleaved uint32 in/out (four channels per thread)"
    return _hbm("pdm2_update on an input stream, 1 Mi channels x 1,024 ticks, " + how, N * F, "samples", ms, 8.0,
                hbm_peak, "4 B in + 4 B out per sample; 4 GiB in + 4 GiB out")


def c1_long(st, ctx, reps=3):
    """The reference's own shape made long: ONE voice of the test_cproc chain, 16 Mi ticks, time-parallel exact scan."""
    rng = np.random.default_rng(9)
    N, F = 1, 16 * 1024 * 1024
    d_in = ctx.dev_alloc(4 * N * F); d_out = ctx.dev_alloc(4 * N * F)
    ctx.h2d(d_in, rng.integers(0, 2, (N, F), dtype=np.uint32))
    b = ctx.batch(st.GRAPH, N, nodes=[(st.NODE_EDGE, -1, 1), (st.NODE_ACC, 0, 1)], mode=1)
    ms = _time(ctx, lambda: b.run_dev(F, inp=d_in, out=d_out), reps)
    b.free(); ctx.dev_free(d_in); ctx.dev_free(d_out)
    return {"config": "C1 long: test_cproc chain, 1 voice x 16 Mi ticks, time-parallel exact scan (CPROC_CUDA_GRAPH_SCAN)", "value": N * F / (ms * 1e-3),
            "unit": "ticks/s", "ms": ms, "bound": "latency / launch (6 small kernels)", "note": "one thread walking the stream: 177 ms"}


def run_all(st, ctx, hbm_peak):
    try:
        cpu = cpu_baselines()
    except Exception as e:
        cpu = {"error": "%s: %s" % (type(e).__name__, e)}
    rows = []
    for key, fn in (("c1", lambda: c1(st, ctx)), ("c2_v1", lambda: c2_v1(st, ctx)), ("c2", lambda: c2_layout(st, ctx, "planar")), ("c2", lambda: c2_layout(st, ctx, "interleaved")),
                    ("c3a", lambda: c3a(st, ctx, hbm_peak, layout="planar")),
                    ("c3a", lambda: c3a(st, ctx, hbm_peak, layout="interleaved")),
                    ("c3b", lambda: c3b(st, ctx)), ("c4", lambda: c4(st, ctx)), ("c4p", lambda: c4p(st, ctx)),
                    ("c5", lambda: c5(st, ctx, hbm_peak, layout="tiled")), ("c5", lambda: c5(st, ctx, hbm_peak, layout="planar")),
                    ("graph_bp5", lambda: graph_bp5(st, ctx, hbm_peak, layout="planar")), ("graph_bp5", lambda: graph_bp5(st, ctx, hbm_peak, layout="interleaved")),
                    ("graph_voice", lambda: graph_voice(st, ctx, hbm_peak, layout="planar")), ("graph_voice", lambda: graph_voice(st, ctx, hbm_peak, layout="interleaved")),
                    ("pdm_raw", lambda: pdm_raw(st, ctx, hbm_peak)), ("pdm_raw", lambda: pdm_raw(st, ctx, hbm_peak, layout="interleaved")), ("c1_long", lambda: c1_long(st, ctx))):
        try:
            row = fn()
            if key in cpu:
                row["cpu_baseline"] = cpu[key]
            elif "error" in cpu:
                row["cpu_baseline"] = cpu
            rows.append(row)
        except Exception as e:                       # never lose the headline line over a secondary row
            rows.append({"error": "%s: %s" % (type(e).__name__, e)})
    return rows


if __name__ == "__main__":
    import json
    import synth_tools_b200 as st
    ctx = st.Context(0)
    for r in run_all(st, ctx, 6538.0):
        print(json.dumps(r))
