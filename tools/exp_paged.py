#!/usr/bin/env python
"""What do paged tables cost?  -m4 (dense tables 385 MiB per block) on 256 x 1 MiB text blocks: dense tables
(one wave, 103 GB), paged with the pool the planner would take, paged with small workspace limits (a smaller
address range to walk).  Kernel times from CUDA events; archives must be equal.

  python tools/exp_paged.py [--level 4] [--blocks 256] [--block-kib 1024]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=4)
    ap.add_argument("--blocks", type=int, default=256)
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--limits-gib", default="0")
    args = ap.parse_args()
    import datagen
    import zpaq_v_b200 as z
    nb, bb = args.blocks, args.block_kib * 1024
    data = datagen.text_stream(nb * bb)
    blocks = [data[i * bb:(i + 1) * bb].tobytes() for i in range(nb)]
    want = data[:nb * bb].tobytes()
    first = None
    runs = [("dense", 1, 0, "1")]
    for g in args.limits_gib.split(","):
        runs += [("paged", 2, int(g), "1"), ("paged", 2, int(g), "0")]
    for name, mode, gib, ahead in runs:
        os.environ["ZPAQGPU_AHEAD"] = ahead  # read when the context is created
        ctx = z.Context(0)
        ctx.set_table_mode(mode)
        ctx.set_workspace_limit(gib << 30)
        res = {"level": args.level, "blocks": nb, "block_kib": args.block_kib, "tables": name, "ws_limit_gib": gib, "ahead": int(ahead)}
        try:
            for rep in range(2):
                arc = ctx.compress_blocks(args.level, blocks)
                st_c = ctx.stats()
                plain, segs, status = ctx.decompress_archive(b"".join(arc))
                st_d = ctx.stats()
                assert status == 0 and plain == want
            if first is None:
                first = arc
            res.update({"equal": arc == first, "enc_ms": round(st_c["codec_ms"], 1), "dec_ms": round(st_d["codec_ms"], 1),
                        "enc_mb_s": round(nb * bb / st_c["codec_ms"] / 1e3, 1), "dec_mb_s": round(nb * bb / st_d["codec_ms"] / 1e3, 1),
                        "waves": [st_c["waves"], st_d["waves"]], "paged": [st_c["paged"], st_d["paged"]],
                        "pool_mb_used": round(st_d["pool_bytes_used"] / 1e6, 1), "init_ms": [round(st_c["init_ms"], 1), round(st_d["init_ms"], 1)]})
        except Exception as ex:
            res["error"] = str(ex)[:200]
        print(json.dumps(res), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
