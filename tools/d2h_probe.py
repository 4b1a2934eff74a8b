"""Concurrent device-to-host copy bandwidth per rank (development tool): is the multi-GPU e2e cap the host's?
torchrun --nproc-per-node N tools/d2h_probe.py"""
import os, time
import torch
import torch.distributed as dist
r = int(os.environ.get("LOCAL_RANK", "0")); w = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(r)
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(r))
except Exception:
    pass
if w > 1:
    dist.init_process_group("gloo")
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.copy_(d); torch.cuda.synchronize()
for label in ("alone" if w == 1 else "concurrent",):
    if w > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("rank %d of %d: D2H %s %.1f GB/s" % (r, w, label, 10 * n / dt / 1e9), flush=True)
