#!/bin/sh
# round 2, eighth GPU pass: AHEAD decoder on the old code generation, A/B, -m4/-m5 at full size
mkdir -p gpurun_out
python -m pytest tests/test_gpu_paged.py tests/test_gpu_determinism.py -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest8.txt; cat gpurun_out/r02_pytest8.txt
python tools/exp_paged.py --level 4 --blocks 256 > gpurun_out/r02_exp_paged4.jsonl 2>&1; cat gpurun_out/r02_exp_paged4.jsonl
python tools/exp_paged.py --level 5 --blocks 296 --block-kib 256 > gpurun_out/r02_exp_paged5.jsonl 2>&1; cat gpurun_out/r02_exp_paged5.jsonl
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 4,5 > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; tail -c 300 gpurun_out/r02_bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench8.json').read().strip().splitlines()[-1])
print("m2", d["compress_mb_s"], d["decompress_mb_s"])
for k,v in d['per_level'].items():
    print(k, json.dumps({x:v[x] for x in v if x in ('compress_kernel_mb_s','decompress_kernel_mb_s','compress_kernel_ms','decompress_kernel_ms','byte_identical_to_oracle')}))
PY
ZPAQGPU_AHEAD=0 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 5 > gpurun_out/r02_bench8b.json 2> gpurun_out/r02_bench8b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench8b.json').read().strip().splitlines()[-1])
for k,v in d['per_level'].items():
    print("AHEAD=0", k, json.dumps({x:v[x] for x in v if x in ('compress_kernel_mb_s','decompress_kernel_mb_s','compress_kernel_ms','decompress_kernel_ms','byte_identical_to_oracle')}))
PY
