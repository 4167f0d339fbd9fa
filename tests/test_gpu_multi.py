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


# ---- several devices behind one C handle (zpaqgpu_multi, csrc/multi.cu) -------------------------------------
def _devices(n):
    """n device indices: distinct GPUs when the box has them, else the same GPU n times (the split, the
    host threads and the two-phase landing are the same code either way)."""
    import torch
    have = torch.cuda.device_count()
    return [k % have for k in range(n)]


@pytest.mark.parametrize("world", [1, 2, 3])
def test_multi_compress_equals_single_and_oracle(gpu_ctx, world):
    import zpaq_v_b200 as z
    m = z.Multi(_devices(world))
    try:
        assert m.device_count() == world
        blocks = [datagen.mixed_block(k, 2000 + 3000 * (k % 5)) for k in range(11)] + [b""]
        names = ["f%d" % k for k in range(len(blocks))]
        comments = ["%d bytes" % len(b) for b in blocks]
        got = m.compress_blocks(2, blocks, names=names, comments=comments)
        assert got == [ob.compress_block(2, b, n, c) for b, n, c in zip(blocks, names, comments)]
        assert got == gpu_ctx.compress_blocks(2, blocks, names=names, comments=comments)
        st = m.stats()
        assert sum(s["n_units"] for s in st) == len(blocks)
        assert [s["first_unit"] for s in st] == sorted(s["first_unit"] for s in st)
        if world > 1:
            assert sum(1 for s in st if s["n_units"] > 0) > 1
        # fewer blocks than devices, and none at all
        assert m.compress_blocks(1, blocks[:1]) == [ob.compress_block(1, blocks[0], "", "")]
        assert m.compress_blocks(1, []) == []
        arc = b"".join(got)
        plain, segs, status = m.decompress_archive(arc)
        want_plain, want_segs, want_status = gpu_ctx.decompress_archive(arc)
        assert (plain, segs, status) == (want_plain, want_segs, want_status)
        assert plain == b"".join(blocks) and status == 0 and all(s["sha1_ok"] == 1 for s in segs)
        assert [s["block_index"] for s in segs] == list(range(len(blocks)))
    finally:
        m.close()


def test_multi_decompress_falls_back_when_a_block_spans_the_cut(gpu_ctx):
    """An archive stored inside an archive: the locators of the inner blocks lie inside the outer block, so a
    cut placed at one of them splits a block.  A damaged block in the middle stops the reference's walk.
    Either way one device repeats the whole archive and the result equals the single-device call's."""
    import zpaq_v_b200 as z
    m = z.Multi(_devices(2))
    try:
        inner = b"".join(ob.compress_block(1, datagen.text(3000, 60 + k), "in%d" % k, "") for k in range(6))
        outer = ob.compress_block(0, inner, "nested.zpaq", "") + ob.compress_block(2, b"after", "t", "")
        got = m.decompress_archive(outer)
        assert got == gpu_ctx.decompress_archive(outer) and got[0] == inner + b"after"
        assert m.stats()[0]["fallback_single"] == 1
        blocks = [ob.compress_block(2, datagen.text(4000, 70 + k), "b%d" % k, "") for k in range(6)]
        bad = bytearray(b"".join(blocks))
        bad[len(blocks[0]) + 16] = 9            # second block: level byte not 1/2, find_block fails there
        got = m.decompress_archive(bytes(bad))
        assert got == gpu_ctx.decompress_archive(bytes(bad)) and got[0] == datagen.text(4000, 70)
    finally:
        m.close()


def test_multi_jidac_add(gpu_ctx):
    """zpaqgpu_multi_jidac_add: one device writes the bytes of zpaqgpu_jidac_add; two devices store a file that
    both ranges hold once, write exactly what the Python host logic (sharding.jidac_add_sharded, two ranks) writes,
    and the archive restores every file."""
    import zpaq_v_b200 as z
    from zpaq_v_b200 import sharding
    a, b = datagen.text(60000, 21), datagen.random_bytes(25000, 22)
    files = {"a.txt": a, "b.bin": b, "c.txt": datagen.text(40000, 23), "empty": b"",
             "a-copy.txt": a, "d.txt": datagen.text(30000, 24), "b-copy.bin": b, "tail": a[:20000] + b[:5000]}
    names, fl = list(files), list(files.values())
    kw = dict(level=1, fragment=2, dedup=True, block_bytes=16384)
    single = gpu_ctx.jidac_add(names, fl, DATE, **kw)
    assert single == ob.jidac_add(names, fl, DATE, **kw)
    one = z.Multi(_devices(1))
    try:
        assert one.jidac_add(names, fl, DATE, **kw) == single
        assert one.jidac_add([], [], DATE, **kw) == gpu_ctx.jidac_add([], [], DATE, **kw)
    finally:
        one.close()
    two = z.Multi(_devices(2))
    try:
        arc = two.jidac_add(names, fl, DATE, **kw)
    finally:
        two.close()
    recs = gpu_ctx.jidac_extract(arc)
    assert {r["name"]: r["data"] for r in recs} == files and all(r["sha1_ok"] == 1 for r in recs)
    # the same archive from the Python host logic with two ranks played one after the other
    def frag(fs):
        return gpu_ctx.jidac_fragment(fs, 2, False)[0]

    def comp(level, blocks, nm, cm):
        return gpu_ctx.compress_blocks(level, blocks, names=nm, comments=cm)

    box = {}
    for _ in range(2):
        arcs = [sharding.jidac_add_sharded(_Rank(2, r, box), frag, comp, names, fl, DATE, level=1, fragment=2,
                                           block_bytes=16384) for r in (1, 0)]
    assert arc == arcs[1]
    # no dedup, one fragment per file, store: the reference's own create_archive bytes through two devices
    two = z.Multi(_devices(2))
    try:
        ref_kw = dict(level=0, fragment=-1, dedup=False, block_bytes=0)
        assert two.jidac_add(names, fl, DATE, **ref_kw) == ob.jidac_add(names, fl, DATE, **ref_kw)
    finally:
        two.close()
