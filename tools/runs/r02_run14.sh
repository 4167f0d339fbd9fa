#!/bin/sh
# equal CTA rounds: -m5 encoder (1 024 blocks, 6 per CTA at most), cfg 5 shape (10 000 files as blocks, -m1, paged)
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 5,jidac > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err; tail -c 300 gpurun_out/r02_bench14.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench14.json').read().strip().splitlines()[-1])
print("m2", d["compress_mb_s"], d["decompress_mb_s"])
for k,v in d['per_level'].items():
    print(k, json.dumps({x:v[x] for x in v if x in ('error','blocks_per_gpu','compress_kernel_mb_s','decompress_kernel_mb_s','warps_per_cta','add_mb_s','byte_identical_to_oracle','byte_identical_to_oracle_on_subtree')}))
PY
python tools/run_configs.py --cfg 5 --files 10000 > gpurun_out/r02_cfg5.json 2>gpurun_out/r02_cfg5.err; cut -c1-700 gpurun_out/r02_cfg5.json
