#!/bin/sh
# bench.py after the table-sizing pass was added to per_level: a short line with -m1 over 160 blocks (more than the
# 148-block warm-up pass, so that the sizing pass runs)
mkdir -p gpurun_out
timeout 80 python bench.py --blocks 160 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 1 > gpurun_out/r02_bench_presize.json 2> gpurun_out/r02_bench_presize.err; echo rc=$?; tail -c 400 gpurun_out/r02_bench_presize.err; python - <<'P'
import json
for l in open('gpurun_out/r02_bench_presize.json'):
    if l.startswith('{'):
        d = json.loads(l); m = d['per_level']['m1']
        print({k: m.get(k) for k in ('compress_mb_s', 'compress_kernel_mb_s', 'decompress_mb_s', 'decompress_kernel_mb_s', 'byte_identical_to_oracle', 'timing', 'presize_error', 'error')})
        print(d['value'], d['compress_mb_s'], d['decompress_mb_s'], d['byte_identical_to_oracle'])
P
