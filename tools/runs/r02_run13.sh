#!/bin/sh
# after the wave-size fix: wave / determinism tests, cfg 3 on one GPU, run_multi at 2 048 blocks on one GPU
mkdir -p gpurun_out
python -m pytest tests/test_gpu_determinism.py tests/test_gpu_more.py tests/test_gpu_paged.py -x -q --durations=8 2>&1 | tail -16
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 3 > gpurun_out/r02_bench13.json 2> gpurun_out/r02_bench13.err; tail -c 300 gpurun_out/r02_bench13.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench13.json').read().strip().splitlines()[-1])
for k,v in d['per_level'].items():
    print(k, json.dumps({x:v[x] for x in v if x in ('error','blocks_per_gpu','compress_kernel_mb_s','decompress_kernel_mb_s','compress_mb_s','decompress_mb_s','waves','issue_frac','byte_identical_to_oracle')}))
PY
