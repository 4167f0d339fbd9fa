#!/bin/sh
# experiment: speculative probe on/off at full occupancy (-m5 1024 x 4 MiB paged, -m4 1024 x 1 MiB paged, -m2 headline)
mkdir -p gpurun_out
for S in 1 0; do
ZPAQGPU_SPEC_PROBE=$S python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 4,5 > gpurun_out/r02_spec$S.json 2> gpurun_out/r02_spec$S.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_spec$S.json').read().strip().splitlines()[-1])
print("spec=$S m2 dec", d["decompress_mb_s"])
for k,v in d['per_level'].items():
    print("spec=$S", k, json.dumps({x:v[x] for x in v if x in ('error','decompress_kernel_mb_s','decompress_kernel_ms','byte_identical_to_oracle')}))
PY
done
