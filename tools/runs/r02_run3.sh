#!/bin/sh
# round 2, third GPU pass: the generic warp kernel (parity + A/B against the one-lane kernel)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_more.py -x -q -k "generic" 2>&1 | tail -25 > gpurun_out/r02_pytest3.txt; cat gpurun_out/r02_pytest3.txt
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level generic > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; tail -c 600 gpurun_out/r02_bench3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench3.json').read().strip().splitlines()[-1])
print(json.dumps(d['per_level']))
PY
