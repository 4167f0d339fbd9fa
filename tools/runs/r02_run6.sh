#!/bin/sh
# round 2, sixth GPU pass: state check after the revert, paging cost experiment, genwarp profile, sanitizers
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-per-level > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench6.json').read().strip().splitlines()[-1])
print("m2", d["compress_mb_s"], d["decompress_mb_s"], d["value"])
PY
python tools/exp_paged.py > gpurun_out/r02_exp_paged.jsonl 2>&1; cat gpurun_out/r02_exp_paged.jsonl
ncu --set full --clock-control none --import-source on -k 'regex:k_decode_genwarp' -c 1 -o gpurun_out/r02_genwarp_dec python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --blocks 64 --block-kib 64 --per-level generic > gpurun_out/r02_ncu_genwarp.log 2>&1
ls -la gpurun_out/r02_genwarp_dec.ncu-rep
sh tools/runs/r02_sanitize.sh
