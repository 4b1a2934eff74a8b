"""Device-to-host copy bandwidth per rank: is the multi-GPU e2e cap the host's?   

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py > profiles/rN_d2h_probe_Ngpu.json

Every rank copies 1 GiB device -> pinned host ten times, (a) one rank at a time, (b) all ranks at once, (c) all at once into a
2 MiB-page-backed buffer (anonymous mmap + MADV_HUGEPAGE + cudaHostRegister), each rank bound to the CPUs next to its GPU (NVML
affinity, like bench.py's e2e leg).  Rank 0 prints ONE JSON line: GB/s per rank for each case, the sum, and where the ranks ran.
bench.py's e2e leg moves 1 byte out per sample: its ceiling per rank is (c)/(b) GB/s = that many G samples/s."""
import ctypes
import json
import mmap
import os
import time

import torch
import torch.distributed as dist

r = int(os.environ.get("LOCAL_RANK", "0")); w = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(r)
cpus = "unbound"
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(r))
    a = sorted(os.sched_getaffinity(0))
    cpus = "%d cpus %d..%d" % (len(a), a[0], a[-1])
except Exception as e:          # no NVML: unbound
    cpus = "unbound (%s)" % type(e).__name__
if w > 1:
    dist.init_process_group("gloo")
n = 1 << 30
REPS = 10
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.copy_(d); torch.cuda.synchronize()


def barrier():
    if w > 1:
        dist.barrier()


def rate(dst):
    t0 = time.perf_counter()
    for _ in range(REPS):
        dst.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return REPS * n / (time.perf_counter() - t0) / 1e9


alone = 0.0
for turn in range(w):                      # (a) one rank at a time
    barrier()
    if turn == r:
        alone = rate(h)
barrier()
conc = rate(h)                             # (b) all ranks at once
# (c) 2 MiB pages: anonymous mapping advised to huge pages, touched, then registered with CUDA
huge = None
try:
    mm = mmap.mmap(-1, n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    mm.madvise(mmap.MADV_HUGEPAGE)
    buf = (ctypes.c_ubyte * n).from_buffer(mm)
    ctypes.memset(buf, 0, n)
    rc = torch.cuda.cudart().cudaHostRegister(ctypes.addressof(buf), n, 0)
    if int(rc) == 0:
        hh = torch.frombuffer(buf, dtype=torch.uint8)
        hh.copy_(d); torch.cuda.synchronize()
        barrier()
        huge = rate(hh)
        torch.cuda.cudart().cudaHostUnregister(ctypes.addressof(buf))
    else:
        barrier()
except Exception:
    barrier()
rec = {"rank": r, "cpus": cpus, "alone_gbs": round(alone, 2), "concurrent_gbs": round(conc, 2), "concurrent_hugepage_gbs": None if huge is None else round(huge, 2)}
if w > 1:
    allr = [None] * w
    dist.all_gather_object(allr, rec)
else:
    allr = [rec]
if r == 0:
    out = {"probe": "device -> pinned host, 1 GiB x %d per rank" % REPS, "n_gpus": w, "ranks": allr,
           "sum_alone_gbs": round(sum(x["alone_gbs"] for x in allr), 1), "sum_concurrent_gbs": round(sum(x["concurrent_gbs"] for x in allr), 1),
           "sum_concurrent_hugepage_gbs": None if any(x["concurrent_hugepage_gbs"] is None for x in allr) else round(sum(x["concurrent_hugepage_gbs"] for x in allr), 1)}
    print(json.dumps(out), flush=True)
if w > 1:
    dist.destroy_process_group()
