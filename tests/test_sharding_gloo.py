"""world_size-2 gloo test of the multi-GPU host logic: disjoint block ranges per
rank, no data-path collective, host concatenation in block order equals the single-process archive.
The per-rank coder is the CPU oracle here (no GPU in this container); on the GPU box the same
functions are driven with Context.compress_blocks and Context.jidac_fragment (tests/test_gpu_multi.py)."""
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


# ---- jidac add over two ranks: the digest exchange ------------------------------------------------
def _oracle_segments(arc):
    """(segments, plaintext) of an archive through the oracle's Decompresser, in the shape
    Context.decompress_archive returns them."""
    import oracle_binding as ob
    d = ob.Decompresser()
    d.set_input(arc)
    segs, plain = [], bytearray()
    while d.find_block():
        while d.find_filename():
            name = d.get_filename()
            before = len(d.output())
            d.decompress(-1)
            d.read_segment_end()
            assert d.last_sha1_ok() == 1
            out = d.output()
            segs.append(dict(filename=name, out_off=before, out_len=len(out) - before))
            plain = out
    return segs, bytes(plain)


def _restore(arc):
    from zpaq_v_b200 import jidac
    segs, plain = _oracle_segments(arc)
    frag, files, dblocks = jidac.parse_index(segs, plain)
    out = {}
    for name, ids in files.items():
        out[name] = b"".join(plain[dblocks[frag[i][0]][0] + frag[i][1]:dblocks[frag[i][0]][0] + frag[i][1] + frag[i][2]]
                             for i in ids)
    return out, frag


def _jidac_files():
    import datagen
    a, b = datagen.text(60000, 21), datagen.random_bytes(25000, 22)
    # the second half repeats files of the first half: duplicates that only the exchange can find
    return {"a.txt": a, "b.bin": b, "c.txt": datagen.text(40000, 23), "empty": b"",
            "a-copy.txt": a, "d.txt": datagen.text(30000, 24), "b-copy.bin": b, "tail": a[:20000] + b[:5000]}


def _jidac_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_binding as ob
    from zpaq_v_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = _jidac_files()
    names, fl = list(files), list(files.values())
    arc = sharding.jidac_add_sharded(
        dist, lambda fs: ob.jidac_fragment(fs, 2, False)[0],
        lambda level, blocks, nm, cm: [ob.compress_block(level, b, n, c) for b, n, c in zip(blocks, nm, cm)],
        names, fl, 20260101120000, level=1, fragment=2, block_bytes=16384)
    if rank == 0:
        q.put(arc)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_jidac_add_dedups_across_ranks():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_jidac_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    arc = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = _jidac_files()
    back, frag = _restore(arc)
    assert back == files and list(back) == list(files)
    # every distinct fragment is stored once although its copies sit on different ranks
    single = ob.jidac_add(list(files), list(files.values()), 20260101120000, level=1, fragment=2, dedup=True,
                          block_bytes=16384)
    back1, frag1 = _restore(single)
    assert back1 == files
    assert len(frag) == len(frag1)
    assert sum(f[2] for f in frag.values()) == sum(f[2] for f in frag1.values()) < sum(map(len, files.values()))
    # same fragments, same ids; only the cut of the d blocks may differ at the rank boundary
    assert {k: v[2:] for k, v in frag.items()} == {k: v[2:] for k, v in frag1.items()}


def test_one_rank_jidac_add_equals_the_single_call():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from zpaq_v_b200 import sharding

    class One:
        def get_world_size(self): return 1
        def get_rank(self): return 0
        def all_gather_object(self, out, obj): out[0] = obj
        def gather_object(self, obj, bucket, dst=0): bucket[0] = obj

    files = _jidac_files()
    names, fl = list(files), list(files.values())
    for fragment, bb in ((2, 16384), (0, 0), (6, 1 << 20)):
        got = sharding.jidac_add_sharded(
            One(), lambda fs: ob.jidac_fragment(fs, fragment, False)[0],
            lambda level, blocks, nm, cm: [ob.compress_block(level, b, n, c) for b, n, c in zip(blocks, nm, cm)],
            names, fl, 20260101120000, level=1, fragment=fragment, block_bytes=bb)
        assert got == ob.jidac_add(names, fl, 20260101120000, level=1, fragment=fragment, dedup=True, block_bytes=bb)
