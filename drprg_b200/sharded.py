"""Read-sharded multi-GPU driver (BASELINE config 3): one process per GPU, the index is
replicated, each rank maps a contiguous shard of the reads, and the packed int32 accumulator
[coverage | locus read counts | scalars] is summed in place with ONE allreduce (NCCL over NVLink via
torch.distributed) before the genotype step, which only the root rank has to run (one VCF per sample; SURVEY §8e:
"S7/S8 on rank 0").  Integer sums => bit-exact for any world size."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_reads: int, world: int, rank: int):
    """rank r maps reads [r*n/G, (r+1)*n/G)."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def sample_shard(items, world: int, rank: int):
    """Batch mode (BASELINE config 5): samples are independent, rank r takes samples r, r + G, r + 2G, ... — replicas
    only, no collective (each rank calls drprg_cuda_map_genotype_batch on its share)."""
    return list(items)[rank::world]


def encode_scalars(total_bases: int, n_reads: int):
    """lo24/hi split keeps every int32 partial sum in range for any realistic shard count."""
    return np.array([total_bases & 0xFFFFFF, total_bases >> 24, n_reads & 0xFFFFFF, n_reads >> 24], np.int32)


def decode_scalars(tail4):
    t = [int(x) for x in tail4]
    return t[0] + (t[1] << 24), t[2] + (t[3] << 24)


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def accum_tensor(index):
    """Zero-copy torch view of the index's device accumulator (flushes the scalars into it)."""
    import torch
    ptr, n = index.accum_device_ptr()
    return torch.as_tensor(_DevArray(ptr, n), device=f"cuda:{index.device}")


def allreduce_accum(index, group=None):
    """In-place sum of the accumulators across ranks (NCCL)."""
    import torch.distributed as dist
    t = accum_tensor(index)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def host_threads_for_rank(rank: int, world: int, cores: int | None = None) -> int:
    """Host threads a rank's library pool should use (DRPRG_THREADS): the root rank formats the VCF and verifies the ML
    paths, the others only feed their GPU, so the root gets what the sibling ranks do not need."""
    import os
    cores = cores or os.cpu_count() or 1
    if world <= 1:
        return cores
    return max(2, cores - 2 * (world - 1)) if rank == 0 else 2


def allreduce_accum_host(accum: np.ndarray, group=None) -> np.ndarray:
    """Host-memory variant (gloo) used by the CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(accum, np.int32).copy())
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
