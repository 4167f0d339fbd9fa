#!/bin/sh
# round 2, fourth GPU pass: generic warp kernel v2 (slot registers, tables in shared memory) + ncu of the -m5 kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_more.py -x -q -k "generic" 2>&1 | tail -5 > gpurun_out/r02_pytest4.txt; cat gpurun_out/r02_pytest4.txt
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level generic > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; tail -c 600 gpurun_out/r02_bench4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench4.json').read().strip().splitlines()[-1])
print(json.dumps(d['per_level']))
PY
B="python bench.py --level 5 --blocks 296 --block-kib 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-level"
$B > gpurun_out/r02_m5_plain.json 2>&1; tail -c 400 gpurun_out/r02_m5_plain.json
ncu --set full --clock-control none --import-source on -k 'regex:k_encode' -s 1 -c 1 -o gpurun_out/r02_m5_enc $B > gpurun_out/r02_ncu_m5_enc.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_decode' -s 1 -c 1 -o gpurun_out/r02_m5_dec $B > gpurun_out/r02_ncu_m5_dec.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
