"""world_size-2 gloo test (CPU) of the read-sharding plumbing: shard bounds cover the reads exactly once
and the packed accumulator (with lo24/hi scalars) sums correctly across ranks."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from drprg_b200 import sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_reads, n_kn, n_loci = 1_000_003, 50, 3
    lo, hi = sharded.shard_bounds(n_reads, world, rank)
    rng = np.random.default_rng(rank)
    cov = rng.integers(0, 70000, size=2 * n_kn + n_loci).astype(np.int32)
    bases = (hi - lo) * 150 + 3_000_000_000 * rank  # forces the hi word to matter
    acc = np.concatenate([cov, sharded.encode_scalars(bases, hi - lo)])
    out = sharded.allreduce_accum_host(acc)
    q.put((rank, lo, hi, cov.astype(np.int64), bases, out.astype(np.int64)))
    dist.barrier()
    dist.destroy_process_group()


def test_accumulator_allreduce_world2():
    from drprg_b200 import sharded
    world, port = 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 1_000_003
    want_cov = res[0][3] + res[1][3]
    want_bases = res[0][4] + res[1][4]
    for r in res:
        out = r[5]
        assert (out[:-4] == want_cov).all()
        assert sharded.decode_scalars(out[-4:]) == (want_bases, 1_000_003)


def test_shard_bounds_partition():
    from drprg_b200 import sharded
    for n in (0, 1, 7, 1000, 30_000_000):
        for g in (1, 2, 4, 8):
            b = [sharded.shard_bounds(n, g, r) for r in range(g)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))


def test_host_threads_per_rank():
    from drprg_b200 import sharded
    assert sharded.host_threads_for_rank(0, 1, 16) == 16
    assert sharded.host_threads_for_rank(0, 8, 32) == 18 and sharded.host_threads_for_rank(3, 8, 32) == 2
    assert sharded.host_threads_for_rank(0, 8, 8) == 2


def test_sample_shard_covers_every_sample_once():
    from drprg_b200 import sharded
    items = [f"s{i}.fq.gz" for i in range(96)]
    for g in (1, 2, 4, 8):
        parts = [sharded.sample_shard(items, g, r) for r in range(g)]
        assert sorted(sum(parts, [])) == sorted(items)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_proportional_subshards_partition():
    """bench.py's host-buffer shards (sized by measured upload bandwidth): contiguous, complete, at least one sub-shard per rank"""
    from drprg_b200 import sharded
    for n, w in ((240, [21.7] * 4 + [33.3] * 4), (240, [1, 1]), (240, [50.0]), (8, [50] + [1] * 7), (17, [3, 2, 2, 9, 1])):
        rs = [sharded.proportional_subshards(n, w, r) for r in range(len(w))]
        assert rs[0].start == 0 and rs[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(rs, rs[1:]))
        assert all(len(r) >= 1 for r in rs)
    rs = [len(sharded.proportional_subshards(240, [21.7] * 4 + [33.3] * 4, r)) for r in range(8)]
    assert rs == [24] * 4 + [36] * 4
