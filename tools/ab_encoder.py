#!/usr/bin/env python
"""A/B of encoder variants (ZPAQGPU_ENC_FLAGS, read when a context is created) on one B200: every variant
compresses the same blocks; the archives must be equal to each other and, on a sample, to the CPU oracle.

  python tools/ab_encoder.py [--level 2] [--blocks 1024] [--block-kib 256] [--variants 1,3] [--reps 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--blocks", type=int, default=1024)
    ap.add_argument("--block-kib", type=int, default=256)
    ap.add_argument("--variants", default="1,3")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import datagen
    import oracle_binding as ob
    import zpaq_v_b200 as z
    nb, bb = args.blocks, args.block_kib * 1024
    data = datagen.text_stream(nb * bb)
    blocks = [data[i * bb:(i + 1) * bb].tobytes() for i in range(nb)]
    comments = ["%d bytes" % bb] * nb
    res = {"level": args.level, "blocks": nb, "block_kib": args.block_kib}
    first = None
    for v in args.variants.split(","):
        os.environ["ZPAQGPU_ENC_FLAGS"] = v
        ctx = z.Context(0)
        best = None
        for _ in range(args.reps):
            arc = ctx.compress_blocks(args.level, blocks, comments=comments)
            ms = ctx.stats()["codec_ms"]
            best = ms if best is None else min(best, ms)
        if first is None:
            first = arc
            for k in (0, nb // 2, nb - 1):
                assert arc[k] == ob.compress_block(args.level, blocks[k], "", comments[k]), "differs from the oracle"
        assert arc == first, "variant %s: archive differs" % v
        res["flags%s_ms" % v] = round(best, 2)
        res["flags%s_mb_s" % v] = round(nb * bb / best / 1e3, 1)
        ctx.close()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
