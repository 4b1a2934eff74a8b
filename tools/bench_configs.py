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


def _time(ctx, fn, reps):
    fn(); ctx.sync()
    best = 1e9
    for _ in range(reps):
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
    ms = _time(ctx, lambda: b.run_dev(F, mix=d_mix), reps)
    b.free(); ctx.dev_free(d_mix)
    return _issue("C4 poly voice (phasor + SVF + AR envelope + pan), 4 Mi voices x 512 frames, float stereo mix", N * F, "voice-samples", ms, 18.0,
                  "SURVEY 8d: 18 algorithmic instr per voice-sample; deterministic fixed-order mix")


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


def mix_bus_scaling(st, ctx, torch, dist, stream, dev, rank, world, reps=40):
    """BASELINE.json config 4 at N > 1: 4 Mi reference voices x 512-frame blocks sharded over the ranks,
    the int32 mix bus formed (a) by NCCL all-reduce + conversion kernel, (b) by the one-kernel
    peer-memory bus overlapped with the next block's render.  Strong scaling (total work fixed).
    Device time, max over ranks."""
    from synth_tools_b200 import shard
    rng = np.random.default_rng(99)
    N, F = 4 * 1024 * 1024, 512
    lo, hi = shard.shard_range(N, rank, world)
    v = np.zeros((hi - lo, 2), np.uint32)
    v[:, 0] = note_incs(rng, hi - lo, 0, 128); v[:, 1] = rng.integers(0, 2**32, hi - lo, dtype=np.uint32)
    b = ctx.batch(st.VOICE_BANK, hi - lo, voices_per_bus=0)
    b.upload_state(v)
    imix = [torch.zeros(F, dtype=torch.int32, device=dev) for _ in range(2)]
    out = [torch.zeros(F, dtype=torch.float32, device=dev) for _ in range(2)]
    bus = shard.connect_bus(st.Bus(ctx, 4096, world, rank))
    k = [0]

    def nccl():
        b.run_dev(F, mix=imix[0].data_ptr())
        dist.all_reduce(imix[0])
        b.mix_to_float(imix[0].data_ptr(), out[0].data_ptr(), F)

    def peer():
        s = k[0] & 1
        k[0] += 1
        bus.wait(s)
        b.run_dev(F, mix=imix[s].data_ptr())
        bus.begin(s, imix[s].data_ptr(), F, out_dev=out[s].data_ptr(), scale=st.Bus.SCALE_SAW)

    res = {}
    for name, fn in (("nccl_allreduce_plus_convert", nccl), ("peer_memory_bus_overlapped", peer)):
        for _ in range(5):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        bus.wait(0); bus.wait(1)
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())
    ok = bus.status() == 0
    bus.destroy(); b.free()
    return {"config": "C4' reference voice bank, 4 Mi voices x 512-frame blocks sharded over %d GPUs (strong scaling), int32 mix bus" % world,
            "n_gpus": world, "bus_ok": ok, "ms_per_block": res,
            "voice_samples_per_s": {kk: N * F / (vv * 1e-3) for kk, vv in res.items()},
            "note": "bit-exactness of both bus forms against the single-device oracle: tools/multi_gpu_mix.py"}


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


def pdm_raw(st, ctx, hbm_peak, reps=3):
    """pdm2_update (stm32f103/pdm.h) on a caller-supplied input stream, planar uint32 in / out."""
    rng = np.random.default_rng(10)
    N, F = 1024 * 1024, 1024
    d_in = ctx.dev_alloc(4 * N * F); d_out = ctx.dev_alloc(4 * N * F)
    chunk = rng.integers(0, 2**32, (16384, F), dtype=np.uint32)
    for k in range(N // 16384):
        ctx.h2d(d_in + k * chunk.nbytes, chunk)
    b = ctx.batch(st.PDM, N, order=2, out_shift=24)
    ms = _time(ctx, lambda: b.run_dev(F, inp=d_in, out=d_out), reps)
    b.free(); ctx.dev_free(d_in); ctx.dev_free(d_out)
    return _hbm("pdm2_update on an input stream, 1 Mi channels x 1,024 ticks, planar uint32 in/out (tensor-TMA staging)", N * F, "samples", ms, 8.0,
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
    rows = []
    for fn in (lambda: c1(st, ctx), lambda: c2_v1(st, ctx), lambda: c3a(st, ctx, hbm_peak, layout="planar"), lambda: c3a(st, ctx, hbm_peak, layout="interleaved"),
               lambda: c3b(st, ctx), lambda: c4(st, ctx), lambda: c4p(st, ctx),
               lambda: c5(st, ctx, hbm_peak, layout="tiled"), lambda: c5(st, ctx, hbm_peak, layout="planar"),
               lambda: graph_bp5(st, ctx, hbm_peak, layout="planar"), lambda: graph_bp5(st, ctx, hbm_peak, layout="interleaved"),
               lambda: pdm_raw(st, ctx, hbm_peak), lambda: c1_long(st, ctx)):
        try:
            rows.append(fn())
        except Exception as e:                       # never lose the headline line over a secondary row
            rows.append({"error": "%s: %s" % (type(e).__name__, e)})
    return rows


if __name__ == "__main__":
    import json
    import synth_tools_b200 as st
    ctx = st.Context(0)
    for r in run_all(st, ctx, 6538.0):
        print(json.dumps(r))
