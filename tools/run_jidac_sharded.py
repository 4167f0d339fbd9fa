#!/usr/bin/env python
"""`jidac add` of the configs[4] tree over N GPUs (one process per GPU, torchrun): every rank cuts and
hashes its byte-balanced file range on its GPU, the fragment digests are all-gathered (the one real
exchange step of this code base), every rank codes the d blocks of the fragments it stores, rank 0
writes the index.  Rank 0 prints one JSON line and checks that the archive restores every file.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      tools/run_jidac_sharded.py --files 4000
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import zpaq_v_b200 as z  # noqa: E402
from run_configs import make  # noqa: E402
from zpaq_v_b200 import sharding  # noqa: E402

DATE = 20260101120000


class One:
    def get_world_size(self): return 1
    def get_rank(self): return 0
    def all_gather_object(self, out, obj): out[0] = obj
    def gather_object(self, obj, bucket, dst=0): bucket[0] = obj
    def barrier(self): pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=4000)
    ap.add_argument("--fragment", type=int, default=6)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--block-kib", type=int, default=1024)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    d = One()
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("gloo")   # host-side object exchange; the GPUs share nothing
        d = dist
    _, files = make(5, args)
    names = ["dir%02d/file%05d" % (k % 37, k) for k in range(len(files))]
    total = sum(map(len, files))
    ctx = z.Context(local)
    ctx.jidac_fragment(files[:4], args.fragment, False)
    ctx.compress_blocks(args.level, files[:2], names=names[:2], comments=["x", "y"])
    d.barrier()
    t0 = time.perf_counter()
    arc = sharding.jidac_add_sharded(
        d, lambda fs: ctx.jidac_fragment(fs, args.fragment, False)[0],
        lambda level, blocks, nm, cm: ctx.compress_blocks(level, blocks, names=nm, comments=cm),
        names, files, DATE, level=args.level, fragment=args.fragment, block_bytes=args.block_kib << 10)
    d.barrier()
    t1 = time.perf_counter()
    if d.get_rank() == 0:
        back = z.jidac.extract(arc, ctx)
        ok = all(back[n] == f for n, f in zip(names, files))
        res = {"what": "jidac add sharded", "n_gpus": world, "files": len(files), "input_bytes": total,
               "archive_bytes": len(arc), "seconds": round(t1 - t0, 3), "add_mb_s": round(total / (t1 - t0) / 1e6, 2),
               "extract_ok": ok, "note": "python list API and object collectives inside the timed region"}
        if world == 1:
            res["equals_single_call"] = arc == ctx.jidac_add(names, files, DATE, level=args.level,
                                                             fragment=args.fragment, dedup=True,
                                                             block_bytes=args.block_kib << 10)
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
