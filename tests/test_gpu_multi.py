"""GPU: the mix-bus exchange entry points on one device (world = 1: push to the own slot, reduce,
fused float scale; the overlapped begin/wait form; the exchange fused into the render launch; argument
checks), and -- when the box has two GPUs -- the world-2 run of tools/multi_gpu_mix.py under torchrun over
NVLink peer memory.  The sharding / reduction logic itself is also covered on CPU with gloo
(tests/test_multi_rank_cpu.py)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(3)


@pytest.fixture(scope="module")
def st():
    import synth_tools_b200 as st_
    return st_


@pytest.fixture(scope="module")
def ctx(st):
    c = st.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_bus_world_1_equals_mix_to_float(st, ctx, oracle, mode):
    N, F = 5000, 512
    v = np.zeros((N, 2), np.uint32)
    v[:, 0] = [oracle.note_to_inc(int(n)) for n in rng.integers(0, 128, N)]
    v[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    va = v.copy()
    want_i, want_f = oracle.voice_bank_run(va, N, N, mode, F)
    b = ctx.batch(st.VOICE_BANK, N, voices_per_bus=0, mode=mode)
    b.upload_state(v)
    d_mix = ctx.dev_alloc(4 * F); d_out = ctx.dev_alloc(4 * F)
    b.run_dev(F, mix=d_mix)
    bus = st.Bus(ctx, F, 1, 0)
    assert len(bus.handle()) == st.abi.lib.cproc_cuda_bus_handle_bytes() == 64
    bus.allreduce(d_mix, F, out_dev=d_out, op=st.Bus.OR if mode else st.Bus.SUM, scale=st.Bus.SCALE_SQUARE if mode else st.Bus.SCALE_SAW)
    mix = np.zeros(F, np.int32); out = np.zeros(F, np.float32)
    ctx.d2h(mix, d_mix); ctx.d2h(out, d_out)
    assert np.array_equal(mix.view(np.uint32), want_i[0].view(np.uint32))
    assert np.array_equal(out.view(np.uint32), want_f[0].view(np.uint32))
    # overlapped form, two slots, consecutive epochs
    d_mix2 = [ctx.dev_alloc(4 * F) for _ in range(2)]; d_out2 = [ctx.dev_alloc(4 * F) for _ in range(2)]
    wants = []
    for k in range(5):
        s = k & 1
        bus.wait(s)
        wants.append(oracle.voice_bank_run(va, N, N, mode, F)[1][0].copy())
        b.run_dev(F, mix=d_mix2[s])
        bus.begin(s, d_mix2[s], F, out_dev=d_out2[s], op=st.Bus.OR if mode else st.Bus.SUM, scale=st.Bus.SCALE_SQUARE if mode else st.Bus.SCALE_SAW)
    bus.wait(0); bus.wait(1)
    ctx.d2h(out, d_out2[0]); assert np.array_equal(out.view(np.uint32), wants[4].view(np.uint32))
    ctx.d2h(out, d_out2[1]); assert np.array_equal(out.view(np.uint32), wants[3].view(np.uint32))
    assert bus.status() == 0
    # grain-mix scale
    g = rng.integers(-5000, 5000, F).astype(np.int32)
    ctx.h2d(d_mix, g)
    bus.allreduce(d_mix, F, out_dev=d_out, scale=st.Bus.SCALE_GRAIN)
    ctx.d2h(out, d_out)
    assert np.array_equal(out, g.astype(np.float32) * np.float32(2.0 ** -7))
    # float bus (op 2): identity at world 1
    fl = rng.uniform(-3, 3, F).astype(np.float32)
    ctx.h2d(d_mix, fl)
    bus.allreduce(d_mix, F, op=st.Bus.FSUM)
    got = np.zeros(F, np.float32); ctx.d2h(got, d_mix)
    assert np.array_equal(got.view(np.uint32), (np.float32(0.0) + fl).view(np.uint32))
    with pytest.raises(st.CprocCudaError):
        bus.allreduce(d_mix, F, out_dev=d_out, op=st.Bus.FSUM, scale=st.Bus.SCALE_SAW)
    bus.destroy(); b.free()
    for p in [d_mix, d_out] + d_mix2 + d_out2:
        ctx.dev_free(p)


def test_bus_errors(st, ctx):
    with pytest.raises(st.CprocCudaError):
        st.Bus(ctx, 16, 0, 0)
    with pytest.raises(st.CprocCudaError):
        st.Bus(ctx, 16, 2, 2)
    with pytest.raises(st.CprocCudaError):
        st.Bus(ctx, 0, 1, 0)
    with pytest.raises(st.CprocCudaError):
        st.Bus(ctx, 16, 17, 0)
    d = ctx.dev_alloc(4096)
    two = st.Bus(ctx, 16, 2, 0)
    with pytest.raises(st.CprocCudaError) as e:            # peers not connected yet
        two.allreduce(d, 16)
    assert e.value.code == st.abi.ESTATE
    two.destroy()
    one = st.Bus(ctx, 16, 1, 0)
    with pytest.raises(st.CprocCudaError):                 # larger than the bus
        one.allreduce(d, 64)
    with pytest.raises(st.CprocCudaError):                 # scale without an output buffer
        one.allreduce(d, 16, scale=st.Bus.SCALE_SAW)
    with pytest.raises(st.CprocCudaError):
        one.begin(2, d, 16)
    one.allreduce(d, 0)                                    # empty exchange is a no-op
    one.destroy()
    ctx.dev_free(d)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("bus_mode", [1, 2])
@pytest.mark.parametrize("N,G,F", [(300000, 0, 512), (5000, 0, 2048), (64 * 40, 64, 64), (100000, 25000, 1500)])
def test_fused_bus_world_1(st, ctx, oracle, mode, bus_mode, N, G, F):
    """The exchange as the tail of the render kernel (cproc_cuda_bus_attach), world = 1: a bus split over tiles
    (atomic tickets, self-cleaning accumulators), several buses, several frame ranges; mode 1 (reduce in the launch)
    and mode 2 (the reduce of block k beside the render of block k+1, the last one in bus_flush); four consecutive
    blocks continue the phases and reuse the four slot parities."""
    v = np.zeros((N, 2), np.uint32)
    v[:, 0] = [oracle.note_to_inc(int(n)) for n in rng.integers(0, 128, N)]
    v[::37, 0] = 0
    v[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    va = v.copy()
    Gv = G or N
    n_bus = (N + Gv - 1) // Gv
    b = ctx.batch(st.VOICE_BANK, N, voices_per_bus=G, mode=mode)
    b.upload_state(v)
    bus = st.Bus(ctx, n_bus * F, 1, 0)
    bus.attach(b, bus_mode)
    K = 6
    d_mix = [ctx.dev_alloc(4 * n_bus * F) for _ in range(K)]; d_out = [ctx.dev_alloc(4 * n_bus * F) for _ in range(K)]
    wants = []
    for k in range(K):
        wants.append(oracle.voice_bank_run(va, N, Gv, mode, F))
        b.run_dev(F, mix=d_mix[k] if k % 3 != 2 else None, out=d_out[k])
    bus.flush()
    ctx.sync()
    assert bus.status() == 0
    mix = np.zeros((n_bus, F), np.int32); out = np.zeros((n_bus, F), np.float32)
    for k in range(K):
        ctx.d2h(out, d_out[k])
        assert np.array_equal(out.view(np.uint32), wants[k][1].view(np.uint32)), k
        if k % 3 != 2:
            ctx.d2h(mix, d_mix[k])
            assert np.array_equal(mix, wants[k][0]), k
    assert np.array_equal(b.download_state(), va)
    bus.detach(b)
    # detached again: the launch writes its own mix
    w = oracle.voice_bank_run(va, N, Gv, mode, F)
    b.run_dev(F, mix=d_mix[0], out=d_out[0]); ctx.sync()
    ctx.d2h(out, d_out[0]); ctx.d2h(mix, d_mix[0])
    assert np.array_equal(out.view(np.uint32), w[1].view(np.uint32)) and np.array_equal(mix, w[0])
    bus.destroy(); b.free()
    for p in d_mix + d_out:
        ctx.dev_free(p)


def test_world_2_mix_bus_under_torchrun():
    """Two ranks over NVLink peer memory: tools/multi_gpu_mix.py under torch.distributed.run compares every bus form
    (NCCL + conversion kernel, the bus kernel, the exchange fused into the render in both modes; voice bank saw / square
    and the float bus of the extension voices) with the single-device oracle, bit for bit."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "tools", "multi_gpu_mix.py"), "--check-only"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    checks = [x for x in rows if "check" in x]
    assert len(checks) >= 3
    for x in checks:
        assert all(v is True for k, v in x.items() if k.endswith("bit_exact") or k.endswith("ok")), x
