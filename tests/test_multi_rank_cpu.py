"""CPU, gloo, world_size 2: the multi-GPU host logic (contiguous bank-aligned
shards + integer mix-bus all-reduce) reproduces the single-device result bit
for bit.  The per-shard render is done by the oracle here (no GPU); on the GPU
box the same logic drives libcproc_cuda (tools/multi_gpu_mix.py under torchrun; bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from synth_tools_b200 import shard
    orc = po.Oracle()
    N, F = 1000, 96
    rng = np.random.default_rng(5)
    v = np.zeros((N, 2), np.uint32)
    tab = np.array([orc.note_to_inc(n) for n in range(128)], np.uint32)
    v[:, 0] = np.where(rng.random(N) < 0.8, tab[rng.integers(0, 128, N)], 0)
    v[:100, 0] = 0x7FFFFFF1                      # loud: the int sum wraps
    v[:, 1] = rng.integers(0, 2**32, N, dtype=np.uint32)
    lo, hi = shard.shard_range(N, rank, world)
    mine = v[lo:hi].copy()
    isum, _ = orc.voice_bank_run(mine, hi - lo, hi - lo, mode, F)
    t = torch.from_numpy(isum.copy())
    shard.allreduce_mix(t, mode)
    full = v.copy()
    want_i, want_f = orc.voice_bank_run(full, N, N, mode, F)
    ok = np.array_equal(t.numpy(), want_i)
    # float scale once, after the reduce (synth.c:180 / :194)
    x = t.numpy()[0]
    f = (x.view(np.uint32).astype(np.float32) if mode == 1 else x.astype(np.float32)) * np.float32(2.0 ** -32)
    ok = ok and np.array_equal(f.view(np.uint32), want_f[0].view(np.uint32))
    q.put((rank, lo, hi, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1])
def test_sharded_mix_allreduce_is_bit_exact(mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert [r[3] for r in res] == [True, True]
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 1000


def test_shard_range_is_a_bank_aligned_partition():
    from synth_tools_b200 import shard
    for n, world, m in [(65536, 8, 3), (65536, 2, 3), (10, 4, 3), (7, 8, 1), (4194304, 8, 1), (100, 3, 64)]:
        edges = [shard.shard_range(n, r, world, m) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for a, b in zip(edges, edges[1:]):
            assert a[1] == b[0]
        for lo, hi in edges:
            assert (lo % m == 0 or lo == n) and (hi % m == 0 or hi == n) and lo <= hi
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 2 * m
