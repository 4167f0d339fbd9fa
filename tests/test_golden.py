"""Committed golden vectors (tests/golden/blocks.json, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them.  GPU (marked): libzpaqgpu reproduces them."""
import hashlib
import json
import os

import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    with open(os.path.join(HERE, "golden", "blocks.json")) as f:
        gold = json.load(f)
    ins = make_golden.inputs()
    for name, meta in gold["inputs"].items():
        assert hashlib.sha1(ins[name]).hexdigest() == meta["sha1"], "input generator drifted: " + name
    return gold, ins


def test_oracle_reproduces_golden():
    gold, ins = load()
    assert len(gold["blocks"]) == 54
    for rec in gold["blocks"]:
        data = ins[rec["input"]]
        arc = ob.compress_block(rec["level"], data, rec["input"], "%d bytes" % len(data))
        assert len(arc) == rec["len"] and hashlib.sha1(arc).hexdigest() == rec["sha1"], rec
        if "hex" in rec:
            assert arc.hex() == rec["hex"]


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [1, 2])
def test_gpu_reproduces_golden(gpu_ctx, kernel):
    gold, ins = load()
    gpu_ctx.set_kernel(kernel)
    gpu_ctx.set_workspace_limit(8 << 30)
    try:
        for level in range(6):
            recs = [r for r in gold["blocks"] if r["level"] == level]
            blocks = [ins[r["input"]] for r in recs]
            got = gpu_ctx.compress_blocks(level, blocks, names=[r["input"] for r in recs],
                                          comments=["%d bytes" % len(b) for b in blocks])
            for r, g in zip(recs, got):
                assert len(g) == r["len"] and hashlib.sha1(g).hexdigest() == r["sha1"], r
            plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
            assert status == 0 and plain == b"".join(blocks)
            assert all(s["sha1_ok"] == 1 for s in segs)
    finally:
        gpu_ctx.set_kernel(0)
        gpu_ctx.set_workspace_limit(0)
