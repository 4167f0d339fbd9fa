#!/usr/bin/env python
"""One process, several B200s behind ONE C handle (zpaqgpu_multi_*, csrc/multi.cu): cfg 3's shape
(-m3, 1 MiB mixed text / random / structured blocks) through zpaqgpu_multi_compress_blocks and
zpaqgpu_multi_decompress_archive with HOST buffers, for 1, 2, .. N devices of the box (strong scaling: the
same blocks every time).  Prints one JSON line per device count: end-to-end MB/s (host clock around the
C-ABI call, copies inside), per-device ranges and kernel times, and whether the archive equals the
one-device archive byte for byte.

  python tools/run_multi.py [--gpus 2] [--level 3] [--blocks 2048] [--block-kib 1024] [--distinct 256]
"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0, help="largest device count (0 = all visible)")
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--blocks", type=int, default=2048)
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--distinct", type=int, default=256)
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    top = args.gpus or have
    nb, bb = args.blocks, args.block_kib * 1024
    distinct = min(nb, args.distinct)
    base = datagen.mixed_stream(distinct, bb)
    data = np.tile(base, (nb + distinct - 1) // distinct)[:nb * bb]
    # pinned host buffers, as a host that cares about PCIe speed would use
    src = torch.from_numpy(data).pin_memory()
    total = nb * bb
    cap = total + total // 2 + 4096 * nb
    arc = torch.empty(cap, dtype=torch.uint8).pin_memory()
    back = torch.empty(total, dtype=torch.uint8).pin_memory()
    off = (C.c_uint64 * (nb + 1))(*[k * bb for k in range(nb + 1)])
    out_off = (C.c_uint64 * (nb + 1))()
    comments = (C.c_char_p * nb)(*[b"%d bytes" % bb for _ in range(nb)])
    segs = (zb.Segment * (nb + 8))()
    L = zb.lib()
    first = None
    counts = [g for g in (1, 2, 4, 8) if g <= top]
    for g in counts:
        m = z.Multi([k % have for k in range(g)])
        try:
            need, nseg = C.c_uint64(0), C.c_int(0)
            res = {"devices": g, "distinct_gpus": min(g, have), "level": args.level, "blocks": nb, "block_bytes": bb,
                   "input_bytes": total}
            for rep in range(2):  # the first pass grows the device buffers
                t0 = time.perf_counter()
                m._check(L.zpaqgpu_multi_compress_blocks(m._h, args.level, src.data_ptr(), off, nb, None, comments,
                                                         arc.data_ptr(), cap, out_off, C.byref(need)))
                t1 = time.perf_counter()
                arc_len = int(out_off[nb])
                rc = L.zpaqgpu_multi_decompress_archive(m._h, arc.data_ptr(), arc_len, back.data_ptr(), total,
                                                        C.byref(need), segs, nb + 8, C.byref(nseg))
                m._check(rc)
                t2 = time.perf_counter()
            st = m.stats()
            assert torch.equal(back, src), "round trip mismatch"
            got = bytes(arc[:arc_len].numpy())
            if first is None:
                first = got
            res.update({
                "compress_mb_s": round(total / (t1 - t0) / 1e6, 2), "decompress_mb_s": round(total / (t2 - t1) / 1e6, 2),
                "round_trip_mb_s": round(total / (t2 - t0) / 1e6, 2), "archive_bytes": arc_len,
                "equals_one_device_archive": got == first, "segments": nseg.value,
                "fallback_single": st[0]["fallback_single"],
                "per_device": [{"device": s["device"], "first_block": s["first_unit"], "blocks": s["n_units"],
                                "stage_ms": round(s["stage_ms"], 1), "fetch_ms": round(s["fetch_ms"], 1),
                                "codec_ms": round(s["stats"]["codec_ms"], 1)} for s in st]})
            print(json.dumps(res), flush=True)
        finally:
            m.close()


if __name__ == "__main__":
    main()
