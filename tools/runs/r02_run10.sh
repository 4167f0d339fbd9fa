#!/bin/sh
# experiment: ncu of the AHEAD serial decoder (-m5, paged, small pool so that ncu's save/restore stays small)
mkdir -p gpurun_out
export ZPAQGPU_WS_LIMIT_MB=24576
B="python bench.py --level 5 --blocks 296 --block-kib 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-level"
ZPAQGPU_DEC_AHEAD=0 $B > gpurun_out/r02_m5_a0.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r02_m5_a0.json').read().strip().splitlines()[-1]); print('ahead0', d['compress_mb_s'], d['decompress_mb_s'], d['stats']['decompress']['codec_ms'], d['stats']['decompress']['paged'])"
ZPAQGPU_DEC_AHEAD=1 $B > gpurun_out/r02_m5_a1.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r02_m5_a1.json').read().strip().splitlines()[-1]); print('ahead1', d['compress_mb_s'], d['decompress_mb_s'], d['stats']['decompress']['codec_ms'], d['stats']['decompress']['paged'])"
ZPAQGPU_DEC_AHEAD=1 ncu --set full --clock-control none --import-source on -k 'regex:k_decode' -s 1 -c 1 -o gpurun_out/r02_m5_dec_ahead $B > gpurun_out/r02_ncu_m5_ahead.log 2>&1
ls -la gpurun_out/r02_m5_dec_ahead.ncu-rep
