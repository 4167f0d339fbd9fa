#!/usr/bin/env python
"""A/B of the decoder variants on one B200: correctness on small inputs first (against the CPU oracle),
then kernel time of every variant on the same archive.

  python tools/ab_decoder.py [--level 2] [--blocks 1024] [--block-kib 256] [--variants tree,tree2]

ZPAQGPU_DECODER is read when a context is created, so every variant gets its own context."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--blocks", type=int, default=1024)
    ap.add_argument("--block-kib", type=int, default=256)
    ap.add_argument("--variants", default="tree,tree2")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--skip-check", action="store_true")
    args = ap.parse_args()
    import datagen
    import oracle_binding as ob
    import zpaq_v_b200 as z

    variants = args.variants.split(",")
    if not args.skip_check:
        small = [b"", b"A", b"Hello World!", datagen.text(5000), datagen.random_bytes(3000), bytes(4096),
                 datagen.text(70000, datagen.SEED0 + 3), b"\xff" * 300]
        for v in variants:
            os.environ["ZPAQGPU_DECODER"] = v.split(":")[0]
            os.environ["ZPAQGPU_GUESS"] = v.split(":")[1] if ":" in v else "1"
            os.environ["ZPAQGPU_PULL"] = v.split(":")[2] if v.count(":") > 1 else "0"
            ctx = z.Context(0)
            ctx.set_workspace_limit(8 << 30)
            for level in (1, 2, 3, 4, 5):
                arc = b"".join(ob.compress_block(level, d, "f%d" % i, "%d bytes" % len(d)) for i, d in enumerate(small))
                t0 = time.time()
                plain, segs, status = ctx.decompress_archive(arc)
                ok = status == 0 and plain == b"".join(small) and all(s["sha1_ok"] == 1 for s in segs)
                print("check %-6s m%d %s (%.2fs, kernel %d, warps/cta %d)" % (
                    v, level, "ok" if ok else "MISMATCH", time.time() - t0, ctx.stats()["kernel"],
                    ctx.stats()["warps_per_cta"]), flush=True)
                if not ok:
                    return 1
            ctx.close()

    nb, bb = args.blocks, args.block_kib * 1024
    data = datagen.text_stream(nb * bb)
    blocks = [data[i * bb:(i + 1) * bb].tobytes() for i in range(nb)]
    os.environ["ZPAQGPU_DECODER"] = variants[0].split(":")[0]
    ctx = z.Context(0)
    arc = b"".join(ctx.compress_blocks(args.level, blocks, comments=["%d bytes" % bb] * nb))
    enc_ms = ctx.stats()["codec_ms"]
    ctx.close()
    want = data[:nb * bb].tobytes()
    res = {"level": args.level, "blocks": nb, "block_kib": args.block_kib, "encode_ms": round(enc_ms, 2)}
    for v in variants:
        os.environ["ZPAQGPU_DECODER"] = v.split(":")[0]
        os.environ["ZPAQGPU_GUESS"] = v.split(":")[1] if ":" in v else "1"
        os.environ["ZPAQGPU_PULL"] = v.split(":")[2] if v.count(":") > 1 else "0"
        ctx = z.Context(0)
        best = None
        for _ in range(args.reps):
            plain, segs, status = ctx.decompress_archive(arc)
            assert status == 0 and plain == want, "variant %s: round trip mismatch" % v
            ms = ctx.stats()["codec_ms"]
            best = ms if best is None else min(best, ms)
        res[v + "_ms"] = round(best, 2)
        res[v + "_mb_s"] = round(nb * bb / best / 1e3, 1)
        res[v + "_warps_per_cta"] = ctx.stats()["warps_per_cta"]
        ctx.close()
    print(json.dumps(res), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
