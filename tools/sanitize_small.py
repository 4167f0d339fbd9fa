#!/usr/bin/env python
"""Small end-to-end pass over every kernel family, sized for compute-sanitizer (memcheck / racecheck slow a
kernel down by one to two orders of magnitude): -m2 tree decoder + three-warp encoder, -m5 serial decoder on
paged tables, the generic warp kernel and its one-lane fallback, store mode, the queued streaming calls and the
jidac front end.  Every result is checked against the CPU oracle, so a sanitizer-clean run is also a correct one.

  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
import oracle_binding as ob  # noqa: E402
import zpaq_v_b200 as z  # noqa: E402
from test_oracle_kats import CUSTOM_HEADERS  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    blocks = [datagen.text(n, 5), datagen.random_bytes(n // 3, 6), datagen.structured(n // 2, 7), b"", b"Hello World!"]
    comments = ["%d bytes" % len(b) for b in blocks]
    ctx = z.Context(0)
    ctx.set_workspace_limit(6 << 30)
    for level, mode in ((2, 0), (1, 0), (5, 2), (0, 0)):
        ctx.set_table_mode(mode)
        got = ctx.compress_blocks(level, blocks, comments=comments)
        assert got == [ob.compress_block(level, b, "", c) for b, c in zip(blocks, comments)], level
        plain, segs, status = ctx.decompress_archive(b"".join(got))
        assert status == 0 and plain == b"".join(blocks) and all(s["sha1_ok"] == 1 for s in segs), level
        print("level %d ok (%s tables)" % (level, "paged" if ctx.stats()["paged"] else "dense"), flush=True)
    ctx.set_table_mode(0)
    for name in ("icm_match_mix2_sse", "forward_refs"):
        hdr = bytes(CUSTOM_HEADERS[name])
        got = ctx.compress_blocks(0, blocks[:3], header=hdr)
        assert got == [ob.compress_block(0, b, "", "", header=hdr) for b in blocks[:3]], name
        plain, segs, status = ctx.decompress_archive(b"".join(got))
        assert status == 0 and plain == b"".join(blocks[:3]), name
        print("generic %s ok" % name, flush=True)
    for k, b in enumerate(blocks[:3]):
        assert ctx.block_begin(level=2) == 0 and ctx.segment_begin("f%d" % k, "") == 0
        assert ctx.segment_write(b) == 0 and ctx.segment_end() == 0 and ctx.block_end_queue() == 0
    assert ctx.flush() == b"".join(ob.compress_block(2, b, "f%d" % k, "") for k, b in enumerate(blocks[:3]))
    print("queued stream ok", flush=True)
    files = {"a": blocks[0], "b": blocks[1], "a2": blocks[0], "e": b""}
    kw = dict(level=1, fragment=0, dedup=True, block_bytes=4096)
    arc = ctx.jidac_add(list(files), list(files.values()), 20260101120000, **kw)
    assert arc == ob.jidac_add(list(files), list(files.values()), 20260101120000, **kw)
    assert z.jidac.extract(arc, ctx) == files
    print("jidac ok", flush=True)
    ctx.close()
    os.environ["ZPAQGPU_GENERIC"] = "lane0"
    ctx = z.Context(0)
    hdr = bytes(CUSTOM_HEADERS["twenty"])
    got = ctx.compress_blocks(0, blocks[:2], header=hdr)
    assert got == [ob.compress_block(0, b, "", "", header=hdr) for b in blocks[:2]]
    print("one-lane generic ok", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
