#!/usr/bin/env python
"""BASELINE.json configs[4]: `jidac add` of a synthetic file tree (sizes log-uniform 1 KiB..1 MiB,
30 % exact duplicates, text/binary 70/30) with rolling-hash fragmentation + SHA-1 dedup at -m1,
through zpaqgpu_jidac_add with HOST buffers.  Prints one JSON line: per-stage kernel times, whole
call MB/s, dedup effect, extraction round trip, byte parity against the CPU oracle on a sub-tree and
the oracle's own single-thread MB/s on that sub-tree.

  python tools/run_jidac.py --files 10000 [--fragment 6] [--level 1] [--block-kib 1024]
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

import oracle_binding as ob  # noqa: E402
import zpaq_v_b200 as z  # noqa: E402
from run_configs import make  # noqa: E402

DATE = 20260101120000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=10000)
    ap.add_argument("--fragment", type=int, default=6)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--oracle-files", type=int, default=200)
    ap.add_argument("--no-extract", action="store_true")
    args = ap.parse_args()
    _, files = make(5, args)
    names = ["dir%02d/file%05d" % (k % 37, k) for k in range(len(files))]
    total = sum(map(len, files))
    kw = dict(level=args.level, fragment=args.fragment, dedup=True, block_bytes=args.block_kib << 10)
    ctx = z.Context(0)
    ctx.jidac_add(names[:4], files[:4], DATE, **kw)        # warm the context
    # the timed call goes straight to the C ABI with preallocated host buffers (no Python copies)
    import ctypes as C
    import numpy as np
    from zpaq_v_b200 import binding as zb
    src = np.frombuffer(b"".join(files), dtype=np.uint8)
    off = np.zeros(len(files) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(f) for f in files])
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    opts = zb.JidacOpts(DATE, kw["level"], kw["fragment"], int(kw["dedup"]), 0, kw["block_bytes"])
    out = np.empty(total + total // 4 + 4096 * (len(files) + 4), dtype=np.uint8)
    ln, need = C.c_uint64(0), C.c_uint64(0)
    t0 = time.perf_counter()
    rc = zb.lib().zpaqgpu_jidac_add(ctx._h, C.byref(opts), arr, src.ctypes.data, off.ctypes.data, len(files),
                                    out.ctypes.data, out.nbytes, C.byref(ln), C.byref(need))
    t1 = time.perf_counter()
    ctx._check(rc)
    arc = out[:ln.value].tobytes()
    st = ctx.jidac_stats()
    res = {"cfg": 5, "what": "jidac add", "files": len(files), "input_bytes": total, "opts": kw,
           "gpu_add_mb_s": round(total / (t1 - t0) / 1e6, 2), "archive_bytes": len(arc),
           "stages_ms": {k: round(st[k], 2) for k in ("h2d_ms", "fragment_ms", "sha1_ms", "dedup_ms", "gather_ms",
                                                      "codec_ms", "pack_ms", "d2h_ms")},
           "fragment_gb_s": round(total / max(st["fragment_ms"], 1e-6) / 1e6, 2),
           "n_fragments": st["n_fragments"], "n_stored": st["n_stored"], "n_dblocks": st["n_dblocks"],
           "stored_bytes": st["stored_bytes"], "launches": st["launches"]}
    if not args.no_extract:
        t2 = time.perf_counter()
        back = z.jidac.extract(arc, ctx)   # zpaqgpu_jidac_extract: decode + fragment gather on the device
        t3 = time.perf_counter()
        res["extract_ok"] = all(back[n] == f for n, f in zip(names, files))
        res["gpu_extract_mb_s"] = round(total / (t3 - t2) / 1e6, 2)
    # parity + CPU baseline on a sub-tree
    k = min(args.oracle_files, len(files))
    sub_total = sum(map(len, files[:k]))
    c0 = time.perf_counter()
    want = ob.jidac_add(names[:k], files[:k], DATE, **kw)
    c1 = time.perf_counter()
    got = ctx.jidac_add(names[:k], files[:k], DATE, **kw)
    res["byte_identical_to_oracle_on_subtree"] = got == want
    res["subtree_files"] = k
    res["cpu_add_mb_s_1thread"] = round(sub_total / (c1 - c0) / 1e6, 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
