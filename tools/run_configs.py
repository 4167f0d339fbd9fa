#!/usr/bin/env python
"""Runs the BASELINE.json configurations (at the sizes given on the command line) once through the
host-buffer C ABI and prints one JSON line per config: GPU compress/decompress MB/s, ratio, waves,
table mode, parity of a sample against the CPU oracle, and the oracle's MB/s on a bounded sample.

  python tools/run_configs.py --cfg 3 --blocks 1024
  python tools/run_configs.py --cfg 4 --blocks 256
  python tools/run_configs.py --cfg 5 --files 10000
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import datagen  # noqa: E402
import oracle_binding as ob  # noqa: E402
import zpaq_v_b200 as z  # noqa: E402


def make(cfg, args):
    if cfg == 2:
        bb = 1 << 20
        whole = datagen.text(args.blocks * bb)
        return 2, [whole[i * bb:(i + 1) * bb] for i in range(args.blocks)]
    if cfg == 3:
        bb = 1 << 20
        return 3, [datagen.mixed_block(k, bb) for k in range(args.blocks)]
    if cfg == 4:
        bb = 4 << 20
        whole = datagen.text(args.blocks * bb, datagen.SEED0 + 4)
        return 5, [whole[i * bb:(i + 1) * bb] for i in range(args.blocks)]
    if cfg == 5:
        files = datagen.file_tree(args.files)[1]
        return 1, files
    raise SystemExit("cfg must be 2..5")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True)
    ap.add_argument("--blocks", type=int, default=256)
    ap.add_argument("--files", type=int, default=10000)
    ap.add_argument("--oracle-blocks", type=int, default=0)
    args = ap.parse_args()
    level, blocks = make(args.cfg, args)
    total = sum(len(b) for b in blocks)
    comments = ["%d bytes" % len(b) for b in blocks]
    ctx = z.Context(0)
    ctx.compress_blocks(level, blocks[:2], comments=comments[:2])   # warm the context
    t0 = time.perf_counter()
    arc = ctx.compress_blocks(level, blocks, comments=comments)
    t1 = time.perf_counter()
    st_c = ctx.stats()
    from zpaq_v_b200 import binding as zb
    note = zb.lib().zpaqgpu_last_error(ctx._h).decode(errors="replace")
    joined = b"".join(arc)
    t2 = time.perf_counter()
    plain, segs, status = ctx.decompress_archive(joined)
    t3 = time.perf_counter()
    st_d = ctx.stats()
    ok = status == 0 and plain == b"".join(blocks) and all(s["sha1_ok"] == 1 for s in segs)
    # parity sample + CPU baseline on a bounded sample
    cores = os.cpu_count() or 1
    n_or = args.oracle_blocks or min(len(blocks), cores * 2)
    idx = list(range(0, len(blocks), max(1, len(blocks) // n_or)))[:n_or]
    import ctypes as C
    L = ob.lib()
    sample = b"".join(blocks[i] for i in idx)
    off = [0]
    for i in idx:
        off.append(off[-1] + len(blocks[i]))
    offs = (C.c_uint64 * len(off))(*off)
    cap = len(sample) + len(sample) // 4 + 4096 * len(idx)
    out = np.empty(cap, dtype=np.uint8)
    out_off = (C.c_uint64 * len(off))()
    need = C.c_uint64(0)
    src = np.frombuffer(sample, dtype=np.uint8)
    c0 = time.perf_counter()
    L.zo_compress_blocks_mt(level, src.ctypes.data, offs, len(idx), out.ctypes.data, cap, out_off, C.byref(need), cores)
    c1 = time.perf_counter()
    back = np.empty(len(sample) + 16, dtype=np.uint8)
    back_off = (C.c_uint64 * len(off))()
    L.zo_decompress_blocks_mt(out.ctypes.data, out_off, len(idx), back.ctypes.data, len(back), back_off, C.byref(need), cores)
    c2 = time.perf_counter()
    identical = all(bytes(out[out_off[k]:out_off[k + 1]]) == arc[i] for k, i in enumerate(idx))
    print(json.dumps({
        "cfg": args.cfg, "level": level, "blocks": len(blocks), "input_bytes": total,
        "gpu_compress_mb_s": round(total / (t1 - t0) / 1e6, 2), "gpu_decompress_mb_s": round(total / (t3 - t2) / 1e6, 2),
        "gpu_compress_kernel_ms": round(st_c["codec_ms"], 1), "gpu_decompress_kernel_ms": round(st_d["codec_ms"], 1),
        "ratio": round(len(joined) / max(total, 1), 4), "roundtrip_ok": ok,
        "byte_identical_to_oracle_on_sample": identical, "sample_blocks": len(idx),
        "waves": [st_c["waves"], st_d["waves"]], "paged": [st_c["paged"], st_d["paged"]],
        "retries": [st_c["retries"], st_d["retries"]], "note": note,
        "pool_mb_used": [round(st_c["pool_bytes_used"] / 1e6, 1), round(st_d["pool_bytes_used"] / 1e6, 1)],
        "cpu_cores": cores, "cpu_compress_mb_s": round(len(sample) / (c1 - c0) / 1e6, 2),
        "cpu_decompress_mb_s": round(len(sample) / (c2 - c1) / 1e6, 2)}))


if __name__ == "__main__":
    main()
