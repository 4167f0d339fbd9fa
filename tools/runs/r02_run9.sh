#!/bin/sh
# round 2, two B200s: the multi-device C handle on real devices, and the bench under torchrun (per_level collectives)
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python tools/run_multi.py --gpus 2 --level 3 --blocks 1024 > gpurun_out/r02_multi2.jsonl 2> gpurun_out/r02_multi2.err; cat gpurun_out/r02_multi2.jsonl | cut -c1-1200; tail -3 gpurun_out/r02_multi2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --cfg3-blocks 2048 --cfg4-blocks 128 --per-level 1,3,5,jidac > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -5 gpurun_out/r02_bench_n2.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"])
for k,v in d['per_level'].items():
    print(k, json.dumps({x:v[x] for x in v if x in ('error','scaling','blocks_per_gpu','compress_kernel_mb_s','decompress_kernel_mb_s','compress_mb_s','decompress_mb_s','byte_identical_to_oracle','add_mb_s')}))
PY
