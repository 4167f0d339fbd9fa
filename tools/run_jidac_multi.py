#!/usr/bin/env python
"""cfg 5 through zpaqgpu_multi_jidac_add: `jidac add` of the synthetic 10 000-file tree (fragment 6, dedup, -m1,
1 MiB d blocks, host buffers) over 1, 2, .. N devices of the box behind one C handle.  One JSON line per
device count: end-to-end MB/s (second call), archive size, extraction check, and whether one device writes the
bytes of zpaqgpu_jidac_add.

  python tools/run_jidac_multi.py [--gpus 2] [--files 10000]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import datagen  # noqa: E402
import zpaq_v_b200 as z  # noqa: E402
from zpaq_v_b200 import binding as zb  # noqa: E402

DATE = 20260101120000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--files", type=int, default=10000)
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    top = args.gpus or have
    names, files = datagen.file_tree(args.files)
    total = sum(map(len, files))
    src = np.frombuffer(b"".join(files), dtype=np.uint8)
    off = np.zeros(len(files) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(f) for f in files])
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    opts = zb.JidacOpts(DATE, 1, 6, 1, 0, 1 << 20)
    out = np.empty(total + total // 4 + 4096 * (len(files) + 4), dtype=np.uint8)
    ctx = z.Context(0)
    single = ctx.jidac_add(names, files, DATE, level=1, fragment=6, dedup=True, block_bytes=1 << 20)
    for g in [k for k in (1, 2, 4, 8) if k <= top]:
        m = z.Multi([k % have for k in range(g)])
        try:
            ln, need = C.c_uint64(0), C.c_uint64(0)
            for rep in range(2):
                t0 = time.perf_counter()
                m._check(zb.lib().zpaqgpu_multi_jidac_add(m._h, C.byref(opts), arr, src.ctypes.data, off.ctypes.data,
                                                          len(files), out.ctypes.data, out.nbytes, C.byref(ln),
                                                          C.byref(need)))
                t1 = time.perf_counter()
            arc = out[:ln.value].tobytes()
            res = {"devices": g, "distinct_gpus": min(g, have), "files": len(files), "input_bytes": total,
                   "add_mb_s": round(total / (t1 - t0) / 1e6, 2), "archive_bytes": len(arc),
                   "per_device": [{"device": s["device"], "files": s["n_units"], "stage_ms": round(s["stage_ms"], 1),
                                   "fetch_ms": round(s["fetch_ms"], 1), "codec_ms": round(s["stats"]["codec_ms"], 1)}
                                  for s in m.stats()]}
            if g == 1:
                res["equals_single_device_call"] = arc == single
            back = z.jidac.extract(arc, ctx)
            res["extract_ok"] = all(back[n] == f for n, f in zip(names, files))
            print(json.dumps(res), flush=True)
        finally:
            m.close()


if __name__ == "__main__":
    main()
