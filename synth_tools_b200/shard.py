"""Multi-GPU host logic: one process per GPU, instances sharded in contiguous
ranges, the only collective being the all-reduce of the shared mix bus.

Every voice / channel / grain is an independent recurrence, so
shards never exchange data-path state.  The mix buses are integer
(synth.c:169-195: int sum / OR of MSBs; square_grain mix in units of 2^-7), so
any split and any reduction tree reproduces the single-device result bit for
bit; the float scale is applied once, after the reduce
(cproc_cuda_mix_to_float).
"""
import torch
import torch.distributed as dist

from . import abi


def shard_range(n_total, rank, world, multiple_of=1):
    """Contiguous [lo, hi) of `n_total` instances for `rank`; boundaries fall on
    multiples of `multiple_of` (dither banks are never split across GPUs)."""
    units = (n_total + multiple_of - 1) // multiple_of
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(lo_u * multiple_of, n_total), min(hi_u * multiple_of, n_total)


def allreduce_mix(imix, mode=abi.MIX_SAW, group=None):
    """In-place all-reduce of an integer mix bus tensor (int32).
    saw / grain mix: wrap-around sum.  square: OR of MSB-only words == max of
    the words viewed as unsigned; NCCL and gloo have no OR, so reduce the sign
    bit as 0/1 with MAX."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return imix
    if mode == abi.MIX_SQUARE:
        bit = (imix != 0).to(torch.int32)
        dist.all_reduce(bit, op=dist.ReduceOp.MAX, group=group)
        imix.copy_(bit * torch.iinfo(torch.int32).min)      # 0 or 0x80000000
        return imix
    # int32 SUM wraps modulo 2^32 on both backends (two's complement adds)
    dist.all_reduce(imix, op=dist.ReduceOp.SUM, group=group)
    return imix


def connect_bus(bus, group=None):
    """Exchange the cudaIpc handles of a cproc_cuda_bus over torch.distributed (any backend)
    and connect it: after this, Bus.allreduce is one kernel per rank over NVLink peer memory."""
    world = dist.get_world_size(group)
    mine = torch.from_numpy(bus.handle().copy())
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    bus.connect(torch.stack([h.cpu() for h in allh]).numpy())
    dist.barrier(group=group)
    return bus
