#!/bin/sh
# round 2, fifth GPU pass: genwarp v3, encoder pull distance A/B, -m4/-m5 with the page-table / MIX2 prefetches
mkdir -p gpurun_out
python -m pytest tests/test_gpu_more.py tests/test_gpu_paged.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest5.txt; cat gpurun_out/r02_pytest5.txt
python tools/ab_encoder.py --variants 1,3,1,3 > gpurun_out/r02_ab_enc.json 2>&1; tail -2 gpurun_out/r02_ab_enc.json
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level generic,4,5 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; tail -c 400 gpurun_out/r02_bench5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench5.json').read().strip().splitlines()[-1])
print("m2", d["compress_mb_s"], d["decompress_mb_s"])
for k,v in d['per_level'].items():
    print(k, json.dumps({x:v[x] for x in v if x not in ('hbm','what','parity_blocks')})[:900])
PY
ZPAQGPU_WS_LIMIT_MB=4096 ncu --set full --clock-control none --import-source on -k 'regex:k_decode_genwarp' -c 1 -o gpurun_out/r02_genwarp_dec python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --blocks 64 --block-kib 64 --per-level generic > gpurun_out/r02_ncu_genwarp.log 2>&1
ls -la gpurun_out/r02_genwarp_dec.ncu-rep
