#!/usr/bin/env python
"""A/B of the SHA-1 kernel: k_sha1_segments (ranges staged through shared memory, the default) against
k_sha1_direct (every lane loads its own range; ZPAQGPU_SHA1=direct).  Each arm is its own process (the choice
is read once per process) and prints one JSON line: SHA-1 milliseconds for 1 MiB segments (the length of the
chain is what the bench configuration has: its 1 024 segments take as long as these 32) on the compress and
the decompress side, and for the fragments of a `jidac add`; plus digests of the outputs, which must agree
between the arms.

  python tools/ab_sha1.py [--blocks 32] [--files 600]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def arm(args):
    import datagen
    import zpaq_v_b200 as z
    bb = 1 << 20
    whole = datagen.text(args.blocks * bb)
    blocks = [whole[i * bb:(i + 1) * bb] for i in range(args.blocks)]
    ctx = z.Context(0)
    ctx.compress_blocks(1, blocks[:2])                                   # warm the context
    arc = ctx.compress_blocks(1, blocks)
    sha_c = ctx.stats()["sha1_ms"]
    plain, segs, status = ctx.decompress_archive(b"".join(arc))
    sha_d = ctx.stats()["sha1_ms"]
    ok = status == 0 and plain == whole and all(s["sha1_ok"] == 1 for s in segs)
    # a tree cut from the same text (datagen.file_tree spends its time on a 256 MiB pool): sizes log-uniform
    # 1 KiB..1 MiB, so the fragments start at every byte alignment
    import numpy as np
    r = datagen._xorshift_stream(datagen.SEED0 + 77, 2 * args.files)
    u = (r[:args.files] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    sizes = np.exp(np.log(1024) + u * (np.log(bb) - np.log(1024))).astype(np.int64)
    files = []
    for k in range(args.files):
        at = int(r[args.files + k] % np.uint64(len(whole) - bb))
        files.append(whole[at:at + int(sizes[k])])
    names = ["dir%02d/file%05d" % (k % 37, k) for k in range(len(files))]
    kw = dict(level=1, fragment=6, dedup=True, block_bytes=bb)
    ctx.jidac_add(names[:4], files[:4], 20260101120000, **kw)
    jarc = ctx.jidac_add(names, files, 20260101120000, **kw)
    js = ctx.jidac_stats()
    print(json.dumps({"sha1": os.environ.get("ZPAQGPU_SHA1", "staged"), "segments": args.blocks, "segment_bytes": bb,
                      "compress_sha1_ms": round(sha_c, 3), "decompress_sha1_ms": round(sha_d, 3), "round_trip_ok": ok,
                      "archive_sha1": hashlib.sha1(b"".join(arc)).hexdigest(),
                      "jidac_files": len(files), "jidac_input_bytes": sum(map(len, files)),
                      "jidac_fragments": js["n_fragments"], "jidac_sha1_ms": round(js["sha1_ms"], 3),
                      "jidac_archive_sha1": hashlib.sha1(jarc).hexdigest()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=32)
    ap.add_argument("--files", type=int, default=600)
    ap.add_argument("--arm", action="store_true")
    args = ap.parse_args()
    if args.arm:
        return arm(args)
    for which in ("staged", "direct"):
        env = dict(os.environ)
        env.pop("ZPAQGPU_SHA1", None)
        if which == "direct":
            env["ZPAQGPU_SHA1"] = "direct"
        subprocess.run([sys.executable, os.path.abspath(__file__), "--arm", "--blocks", str(args.blocks),
                        "--files", str(args.files)], env=env, check=False)


if __name__ == "__main__":
    main()
