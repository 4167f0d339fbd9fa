#!/bin/sh
# round 2, second GPU pass: the new multi-device / queued-stream tests, cfg 3 and -m4 per-level passes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_stream_batch.py tests/test_gpu_multi.py tests/test_gpu_paged.py -x -q -s 2>&1 | tail -25 > gpurun_out/r02_pytest2.txt; cat gpurun_out/r02_pytest2.txt
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 3,4 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; tail -c 600 gpurun_out/r02_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench2.json').read().strip().splitlines()[-1])
for k,v in d['per_level'].items():
    print(k, json.dumps({x:v[x] for x in v if x not in ('hbm','what','parity_blocks')}))
PY
