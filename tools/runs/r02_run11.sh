#!/bin/sh
# experiment: AHEAD v2 serial decoder (shared per-byte hash table, L1 pulls of page-table entries and MIX2 windows)
mkdir -p gpurun_out
ZPAQGPU_DEC_AHEAD=1 python -m pytest tests/test_gpu_paged.py tests/test_gpu_fullsize.py -x -q -k "paged" 2>&1 | tail -3
export ZPAQGPU_WS_LIMIT_MB=65536
for L in "5 296 256" "4 256 1024"; do
set -- $L
B="python bench.py --level $1 --blocks $2 --block-kib $3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-level"
for A in 0 1 0 1; do
ZPAQGPU_DEC_AHEAD=$A $B > gpurun_out/r02_ab_tmp.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r02_ab_tmp.json').read().strip().splitlines()[-1]); print('m$1 ahead$A dec_ms', round(d['stats']['decompress']['codec_ms'],1), 'paged', d['stats']['decompress']['paged'], 'parity', d['byte_identical_to_oracle'])"
done
done
unset ZPAQGPU_WS_LIMIT_MB
ZPAQGPU_DEC_AHEAD=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --per-level 4,5 > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench11.json').read().strip().splitlines()[-1])
for k,v in d['per_level'].items():
    print("AHEAD=1", k, json.dumps({x:v[x] for x in v if x in ('error','compress_kernel_mb_s','decompress_kernel_mb_s','decompress_kernel_ms','byte_identical_to_oracle')}))
PY
