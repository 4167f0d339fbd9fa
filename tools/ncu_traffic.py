#!/usr/bin/env python
"""DRAM traffic per launch of the codec kernels on the bench configuration.

Run on the GPU box (single-pass metrics, so the 1-2 s kernels are not replayed):

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:'k_(en|de)code' --csv --log-file gpurun_out/r02_traffic_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-per-level

then here:  python tools/ncu_traffic.py gpurun_out/r02_traffic_launches.csv --level 2 --blocks 1024 --block-kib 1024
which copies the launch list to profiles/r02_traffic_launches.csv and writes profiles/r02_traffic.json
(the file bench.py reads `roofline.traffic` from)."""
import argparse
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    u = unit.strip().lower()
    mult = {"byte": 1, "bytes": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
    return v * mult.get(u, 1)


def to_ms(value, unit):
    v = float(value.replace(",", ""))
    u = unit.strip().lower()
    return v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3,
                "second": 1e3}.get(u, 1e-6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--blocks", type=int, default=1024)
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--tag", default="r02")
    args = ap.parse_args()
    rows = []
    with open(args.csv, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    launches = {}
    for r in rows:
        key = (r["ID"], r["Kernel Name"])
        d = launches.setdefault(key, {})
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            d[m] = to_bytes(r["Metric Value"], r["Metric Unit"])
        elif m == "gpu__time_duration.sum":
            d[m] = to_ms(r["Metric Value"], r["Metric Unit"])
    out = {}
    for (lid, name), d in launches.items():
        kind = "decode" if "k_decode" in name else ("encode" if "k_encode" in name else None)
        if not kind or "dram__bytes_read.sum" not in d:
            continue
        rec = {"kernel": name.split("(")[0], "launch_id": int(lid), "level": args.level, "blocks": args.blocks,
               "block_bytes": args.block_kib * 1024,
               "dram_bytes_read": int(d["dram__bytes_read.sum"]), "dram_bytes_write": int(d["dram__bytes_write.sum"]),
               "dram_bytes_per_launch": int(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]),
               "ncu_duration_ms": round(d.get("gpu__time_duration.sum", 0.0), 3)}
        rec["dram_bytes_per_input_byte"] = round(rec["dram_bytes_per_launch"] / (args.blocks * args.block_kib * 1024), 2)
        # keep the last full-size launch of each kind (earlier ones are warm-up passes of the same shape)
        if kind not in out or rec["ncu_duration_ms"] >= 0.5 * out[kind]["ncu_duration_ms"]:
            out[kind] = rec
    if not out:
        sys.exit("no codec kernel launches with dram metrics in " + args.csv)
    out["source"] = "profiles/%s_traffic_launches.csv" % args.tag
    dst = os.path.join(ROOT, "profiles", "%s_traffic_launches.csv" % args.tag)
    if os.path.abspath(args.csv) != dst:
        shutil.copyfile(args.csv, dst)
    with open(os.path.join(ROOT, "profiles", "%s_traffic.json" % args.tag), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
