#!/bin/sh
# round 2, seventh GPU pass: AHEAD / PAGED kernel variants (entries in registers): parity, A/B, -m4/-m5 at full size
mkdir -p gpurun_out
python -m pytest tests/test_gpu_paged.py tests/test_gpu_determinism.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest7.txt; cat gpurun_out/r02_pytest7.txt
python tools/exp_paged.py --level 4 --blocks 256 > gpurun_out/r02_exp_paged2.jsonl 2>&1; cat gpurun_out/r02_exp_paged2.jsonl
python tools/exp_paged.py --level 5 --blocks 296 --block-kib 256 > gpurun_out/r02_exp_paged3.jsonl 2>&1; cat gpurun_out/r02_exp_paged3.jsonl
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 4,5 > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; tail -c 300 gpurun_out/r02_bench7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench7.json').read().strip().splitlines()[-1])
print("m2", d["compress_mb_s"], d["decompress_mb_s"])
for k,v in d['per_level'].items():
    print(k, json.dumps({x:v[x] for x in v if x not in ('hbm','what','parity_blocks')})[:700])
PY
