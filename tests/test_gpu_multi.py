"""The multi-rank host logic (zpaq_v_b200.sharding) driven with the real GPU coder instead of the oracle
stand-in of tests/test_sharding_gloo.py.  One process, one GPU: the ranks are played one after the other
through the same Context, which is what every rank does on its own device under torchrun
(tools/run_jidac_sharded.py, bench.py --gpus N)."""
import pytest

import datagen
import oracle_binding as ob

pytestmark = pytest.mark.gpu
DATE = 20260101120000


class _Rank:
    """torch.distributed look-alike for rank r of `world` ranks played sequentially: collectives read
    and write a shared mailbox, so every rank must be run in order and twice (gather after compute)."""

    def __init__(self, world, rank, box):
        self.world, self.rank, self.box = world, rank, box

    def get_world_size(self):
        return self.world

    def get_rank(self):
        return self.rank

    def all_gather_object(self, out, obj):
        self.box.setdefault("ag", {})[self.rank] = obj
        for r in range(self.world):
            out[r] = self.box["ag"].get(r, [])   # empty until that rank has run once (first pass)

    def gather_object(self, obj, bucket, dst=0):
        self.box.setdefault("g", {})[self.rank] = obj
        if bucket is not None:
            for r in range(self.world):
                bucket[r] = self.box["g"].get(r)


def test_block_sharding_with_the_gpu_coder(gpu_ctx):
    from zpaq_v_b200 import sharding
    blocks = [datagen.mixed_block(k, 3000 + 500 * k) for k in range(7)]
    names = ["f%d" % k for k in range(7)]
    comments = ["%d bytes" % len(b) for b in blocks]
    bounds = sharding.shard_by_bytes([len(b) for b in blocks], 2)
    parts = []
    for rank in range(2):
        first, coded = sharding.compress_sharded(
            lambda bl, nm, cm: gpu_ctx.compress_blocks(2, bl, names=nm, comments=cm), blocks, names, comments, 2, rank,
            bounds)
        parts.append((first, coded))
    merged = [b for _, coded in sorted(parts) for b in coded]
    assert merged == [ob.compress_block(2, b, n, c) for b, n, c in zip(blocks, names, comments)]


def test_jidac_add_two_ranks_with_the_gpu_coder(gpu_ctx):
    """Rank 1 holds copies of rank 0's files: they must be stored once, and the archive must restore."""
    import zpaq_v_b200 as z
    from zpaq_v_b200 import sharding
    a, b = datagen.text(60000, 21), datagen.random_bytes(25000, 22)
    files = {"a.txt": a, "b.bin": b, "c.txt": datagen.text(40000, 23), "empty": b"",
             "a-copy.txt": a, "d.txt": datagen.text(30000, 24), "b-copy.bin": b, "tail": a[:20000] + b[:5000]}
    names, fl = list(files), list(files.values())

    def frag(fs):
        return gpu_ctx.jidac_fragment(fs, 2, False)[0]

    def comp(level, blocks, nm, cm):
        return gpu_ctx.compress_blocks(level, blocks, names=nm, comments=cm)

    box = {}
    # pass 1 fills the mailbox with every rank's digests, pass 2 runs with complete tables
    for _ in range(2):
        arcs = [sharding.jidac_add_sharded(_Rank(2, r, box), frag, comp, names, fl, DATE, level=1, fragment=2,
                                           block_bytes=16384) for r in (1, 0)]
    arc = arcs[1]
    assert arcs[0] is None and arc is not None
    recs = gpu_ctx.jidac_extract(arc)
    assert {r["name"]: r["data"] for r in recs} == files and all(r["sha1_ok"] == 1 for r in recs)
    single = gpu_ctx.jidac_add(names, fl, DATE, level=1, fragment=2, dedup=True, block_bytes=16384)
    assert z.jidac.extract(single, gpu_ctx) == files
    # one rank: the sharded path writes the bytes of the single call
    one = sharding.jidac_add_sharded(_Rank(1, 0, {}), frag, comp, names, fl, DATE, level=1, fragment=2,
                                     block_bytes=16384)
    assert one == single == ob.jidac_add(names, fl, DATE, level=1, fragment=2, dedup=True, block_bytes=16384)
