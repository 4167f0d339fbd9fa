"""world_size-2 gloo test of the multi-GPU host logic: disjoint block ranges per
rank, no data-path collective, host concatenation in block order equals the single-process archive.
The per-rank coder is the CPU oracle here (no GPU in this container); on the GPU box the same
functions are driven with Context.compress_blocks (tests/test_gpu_multi.py)."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import datagen
    import oracle_binding as ob
    from zpaq_v_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blocks = [datagen.mixed_block(k, 3000 + 500 * k) for k in range(7)]
    names = ["f%d" % k for k in range(7)]
    comments = ["%d bytes" % len(b) for b in blocks]

    def coder(bl, nm, cm):
        return [ob.compress_block(2, b, n, c) for b, n, c in zip(bl, nm, cm)]

    bounds = sharding.shard_by_bytes([len(b) for b in blocks], world)
    first, parts = sharding.compress_sharded(coder, blocks, names, comments, world, rank, bounds)
    merged = sharding.gather_in_block_order(dist, first, parts)
    if rank == 0:
        want = coder(blocks, names, comments)
        q.put((merged == want, bounds, len(parts)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_concatenates_in_block_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, bounds, n0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert bounds[0] == 0 and bounds[-1] == 7 and 0 < bounds[1] < 7 and n0 == bounds[1]


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    from zpaq_v_b200 import sharding
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 4, 8):
            cuts = [sharding.shard_range(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
    b = sharding.shard_by_bytes([10, 10, 10, 1000, 10, 10], 4)
    assert b[0] == 0 and b[-1] == 6 and len(b) == 5 and b == sorted(b)
