#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep) into the text summaries committed under profiles/.

  python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r01_xyz

writes <out>_raw.txt (per-kernel headline counters) and <out>_stalls_<kernel>.txt (stall reasons
and the hottest SASS instructions from the source page; kernels must be built with -lineinfo)."""
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_issued.avg.per_cycle_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
       "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    head, units = rows[0], rows[1]
    names = []
    with open(out + "_raw.txt", "w") as f:
        for r in rows[2:]:
            k = r[head.index("Kernel Name")]
            names.append(k)
            f.write("kernel: %s\n" % k)
            for m in RAW:
                if m in head:
                    f.write("  %-58s %s %s\n" % (m, r[head.index(m)], units[head.index(m)]))
            f.write("\n")
    for k in names:
        short = "encode" if "encode" in k else "decode" if "decode" in k else "kernel"
        rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name",
                                                 "regex:" + re.search(r"k_[a-z0-9_]+", k).group(0)]))))
        head = rows[1]
        ia, isrc, isamp, iex = (head.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed"))
        stalls = [c for c in head if c.startswith("stall_") and "Not Issued" not in c]
        seen, data = set(), []
        for r in rows[2:]:
            if len(r) >= len(head) and r[ia].startswith("0x") and r[ia] not in seen:
                seen.add(r[ia])
                data.append(r)
        tot = sum(int(r[isamp]) for r in data) or 1
        base = int(data[0][ia], 16)
        with open("%s_stalls_%s.txt" % (out, short), "w") as f:
            f.write("kernel: %s\nwarp-state samples: %d over %d SASS instructions\n\nstall reasons (share of samples)\n"
                    % (k, tot, len(data)))
            agg = {c: sum(int(r[head.index(c)]) for r in data) for c in stalls}
            for c, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
                f.write("  %-28s %6.2f%%\n" % (c, 100.0 * v / tot))
            f.write("\nhottest instructions (offset, share, executed, dominant stall, SASS)\n")
            top = sorted(data, key=lambda r: -int(r[isamp]))[:25]
            for r in sorted(top, key=lambda r: int(r[ia], 16)):
                dom = max(stalls, key=lambda c: int(r[head.index(c)]))
                f.write("  %05x %6.2f%% %12s %-18s %s\n" % (int(r[ia], 16) - base, 100.0 * int(r[isamp]) / tot, r[iex],
                                                            dom[6:], r[isrc].strip()[:70]))


if __name__ == "__main__":
    main()
