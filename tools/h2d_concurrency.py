"""Why does the host-buffer (`e2e`) step stop scaling at 4-8 GPUs?  Under torchrun, every rank uploads a pinned buffer to
its GPU, first one rank at a time, then all ranks at once; with and without the NUMA binding bench.py applies.  Prints the
box topology, where each rank's pinned pages live, and the per-rank / aggregate GB/s.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/h2d_concurrency.py"""
import ctypes, json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from drprg_b200 import sharded

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
MB = 256


def node_of(ptr):
    """NUMA node of the page at ptr (get_mempolicy with MPOL_F_NODE | MPOL_F_ADDR), -1 if unknown"""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        node = ctypes.c_int(-1)
        r = libc.syscall(239, ctypes.byref(node), None, ctypes.c_ulong(0), ctypes.c_void_p(ptr), ctypes.c_ulong(3))
        return node.value if r == 0 else -1
    except Exception:
        return -1


def run(bind):
    info = sharded.bind_to_gpu_numa_node(local) if bind else {"numa_node": None, "cpus": None}
    h = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()

    def copy_ms(reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for _ in range(reps):
                d.copy_(h, non_blocking=True)
        s.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps

    copy_ms(2)
    solo = None
    for r in range(world):           # one rank at a time
        dist.barrier()
        if r == rank:
            solo = copy_ms()
    dist.barrier()
    together = copy_ms(8)            # all at once
    dist.barrier()
    out = {"rank": rank, "bind": bind, "gpu_numa": info.get("numa_node"), "cpus_after_bind": len(os.sched_getaffinity(0)),
           "pinned_on_node": node_of(h.data_ptr()), "solo_GBps": round(MB / 1024 / (solo * 1e-3), 1),
           "concurrent_GBps": round(MB / 1024 / (together * 1e-3), 1)}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        for g in gathered:
            print(json.dumps(g))
        print(json.dumps({"bind": bind, "aggregate_concurrent_GBps": round(sum(g["concurrent_GBps"] for g in gathered), 1)}), flush=True)


torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "-H"]):
        try:
            o = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            print("\n".join(l for l in o.splitlines() if not cmd[0] == "lscpu" or any(k in l for k in ("NUMA", "Model name", "Socket", "CPU(s):"))), flush=True)
        except Exception as e:
            print(cmd[0], "unavailable:", e, flush=True)
full = os.sched_getaffinity(0)
run(False)
os.sched_setaffinity(0, full)
run(True)
dist.barrier()
dist.destroy_process_group()
