"""Read-sharded multi-GPU driver (BASELINE config 3), one process per GPU: the index is replicated, each
rank maps a contiguous shard of the reads, and the packed int32 accumulator [coverage | locus read counts | scalars] is
combined on the root rank, which runs the genotype step (one VCF per sample; SURVEY §8e: "S7/S8 on rank 0").
Two ways to combine, both integer sums and therefore bit-exact for any world size:

  fused (default)   the root exports its accumulator through CUDA IPC (`setup_fused_reduce`); every other rank's coverage
                    merge kernel adds straight into it over NVLink (red.global.add on peer-mapped memory) and signals its
                    arrival on the device; the root's genotype step waits for the arrivals on the device.  No collective
                    call, no host barrier inside a sample.
  allreduce         ONE in-place NCCL allreduce of the accumulator through torch.distributed (`allreduce_accum`): the
                    plain-library form, kept as the cross-check (tools/sharded_parity.py compares the two).

The same sharding inside ONE process is `lib.Index(prg, n_gpus=N)` (drprg_cuda_index_load_multi)."""
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


def setup_fused_reduce(index, rank: int, world: int, group=None):
    """Exchange the root's 64-byte CUDA IPC handle (rank 0 -> everyone) and attach it on the other ranks.  From then on
    `index.map_batch` on a non-root rank adds into the root's accumulator; call `index.shard_done()` after the rank's last
    batch of a sample; the root's `index.genotype()` waits for world - 1 arrivals on the device."""
    import torch.distributed as dist
    if world <= 1:
        return
    box = [index.shard_root(world) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    if rank != 0:
        index.shard_attach(world, box[0])
    dist.barrier(group=group)  # nobody starts a sample before every rank is attached


def host_threads_for_rank(rank: int, world: int, cores: int | None = None) -> int:
    """Host threads a rank's library pool should use (DRPRG_THREADS): the root rank formats the VCF and verifies the ML
    paths, the others only feed their GPU, so the root gets what the sibling ranks do not need."""
    import os
    cores = cores or os.cpu_count() or 1
    if world <= 1:
        return cores
    return max(2, cores - 2 * (world - 1)) if rank == 0 else 2


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the host cores of the NUMA node the GPU hangs off (so that pinned buffers are first-touched
    there and the upload threads run next to the GPU's PCIe root).  A no-op on single-node hosts; never fatal."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        import subprocess
        bus = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]  # sysfs uses a 4-digit domain
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        if node >= 0 and len(nodes) > 1:
            cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["cpus"] = len(cpus)
    except Exception:
        pass
    return info


def allreduce_accum_host(accum: np.ndarray, group=None) -> np.ndarray:
    """Host-memory variant (gloo) used by the CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(accum, np.int32).copy())
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


def concurrent_h2d_gbps(torch, dist, mb: int = 64, reps: int = 6):
    """every rank uploads a pinned buffer to its GPU at the same time: the GB/s each one gets (list over ranks, identical
    on all ranks).  Used to size the host-buffer shards: a sharded step ends when the slowest upload does."""
    import time
    world = dist.get_world_size()
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    best = 0.0
    for rnd in range(3):  # the first round warms up; the better of the next two counts
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        gbps = reps * mb / 1024.0 / (time.perf_counter() - t0)
        if rnd:
            best = max(best, gbps)
    t = torch.zeros(world, dtype=torch.float64, device="cuda")
    t[dist.get_rank()] = best
    dist.all_reduce(t)
    return [float(x) for x in t.cpu()]


def proportional_subshards(n_subshards: int, weights, rank: int) -> range:
    """contiguous ranges of sub-shards, sizes proportional to `weights` (largest-remainder apportionment, at least one
    per rank); returns this rank's range"""
    w = [max(1e-9, float(x)) for x in weights]
    tot = sum(w)
    quota = [n_subshards * x / tot for x in w]
    cnt = [max(1, int(q)) for q in quota]
    while sum(cnt) > n_subshards:  # the minimum of one pushed the sum over: take from the largest
        cnt[cnt.index(max(cnt))] -= 1
    rem = sorted(range(len(w)), key=lambda i: quota[i] - int(quota[i]), reverse=True)
    i = 0
    while sum(cnt) < n_subshards:
        cnt[rem[i % len(rem)]] += 1
        i += 1
    start = sum(cnt[:rank])
    return range(start, start + cnt[rank])
