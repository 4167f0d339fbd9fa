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


def load_jidac():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    with open(os.path.join(HERE, "golden", "jidac.json")) as f:
        gold = json.load(f)
    return gold, make_golden


def _cuts_sha1(frs):
    return hashlib.sha1(b"".join(b"%d,%d,%d;" % (f["off"], f["len"], f["id"]) for f in frs)).hexdigest()


def test_oracle_reproduces_jidac_golden():
    gold, mg = load_jidac()
    names, files = mg.jidac_tree()
    tiny = {"a": b"hello world", "empty": b""}
    assert ob.jidac_add(list(tiny), list(tiny.values()), gold["date"]).hex() == gold["tiny_reference_hex"]
    assert len(gold["archives"]) == len(mg.JIDAC_OPTS)
    for rec in gold["archives"]:
        arc = ob.jidac_add(names, files, gold["date"], **rec["opts"])
        assert len(arc) == rec["len"] and hashlib.sha1(arc).hexdigest() == rec["sha1"], rec["opts"]
        frs, stored = ob.jidac_fragment(files, rec["opts"]["fragment"], rec["opts"]["dedup"])
        assert (len(frs), stored, _cuts_sha1(frs)) == (rec["n_fragments"], rec["n_stored"], rec["cuts_sha1"])


@pytest.mark.gpu
def test_gpu_reproduces_jidac_golden(gpu_ctx):
    gold, mg = load_jidac()
    names, files = mg.jidac_tree()
    tiny = {"a": b"hello world", "empty": b""}
    assert gpu_ctx.jidac_add(list(tiny), list(tiny.values()), gold["date"]).hex() == gold["tiny_reference_hex"]
    gpu_ctx.set_workspace_limit(8 << 30)
    try:
        for rec in gold["archives"]:
            arc = gpu_ctx.jidac_add(names, files, gold["date"], **rec["opts"])
            assert len(arc) == rec["len"] and hashlib.sha1(arc).hexdigest() == rec["sha1"], rec["opts"]
            frs, stored = gpu_ctx.jidac_fragment(files, rec["opts"]["fragment"], rec["opts"]["dedup"])
            assert (len(frs), stored, _cuts_sha1(frs)) == (rec["n_fragments"], rec["n_stored"], rec["cuts_sha1"])
            back = {r["name"]: r["data"] for r in gpu_ctx.jidac_extract(arc)}
            assert back == dict(zip(names, files))
    finally:
        gpu_ctx.set_workspace_limit(0)
