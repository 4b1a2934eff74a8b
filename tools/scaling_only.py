"""The strong-scaling rows of bench.py alone (C4', C4, C3b, C5 sharded over the GPUs of the box, parity checks included), for quick
multi-GPU measurements:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scaling_only.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import synth_tools_b200 as st
from tools import bench_configs

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = st.Context(local, stream.cuda_stream)
rows = bench_configs.scaling_rows(st, ctx, torch, stream, dev, rank, world, dist if world > 1 else None)
if rank == 0:
    print(json.dumps({"n_gpus": world, "scaling_configs": rows}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
