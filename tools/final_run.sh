#!/bin/sh
# Round-end measurement pass on one B200: tests, smoke, bench (both arms), the ncu launch list of the bench
# command itself (one metric, one pass) and full captures of the two codec kernels; with FULL=1 also the
# other BASELINE.json configs.  Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json
python bench.py --impl reference > gpurun_out/bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.json 2>/dev/null
B="python bench.py --blocks 1024 --block-kib 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k 'regex:k_encode' -s 1 -c 1 -o gpurun_out/prof_final_enc $B > gpurun_out/ncu_final_enc.log 2>&1
if [ -n "$FULL" ]; then
ncu --set full --clock-control none --import-source on -k 'regex:k_decode' -s 1 -c 1 -o gpurun_out/prof_final_dec $B > gpurun_out/ncu_final_dec.log 2>&1
python tools/run_configs.py --cfg 3 --blocks 1024 > gpurun_out/cfg3.json 2>gpurun_out/cfg3.err
python tools/run_configs.py --cfg 4 --blocks 256 > gpurun_out/cfg4.json 2>gpurun_out/cfg4.err
python tools/run_configs.py --cfg 5 --files 10000 > gpurun_out/cfg5.json 2>gpurun_out/cfg5.err
python tools/run_jidac.py --files 10000 > gpurun_out/cfg5_jidac.json 2>gpurun_out/cfg5_jidac.err
cat gpurun_out/cfg3.json gpurun_out/cfg4.json gpurun_out/cfg5.json gpurun_out/cfg5_jidac.json | cut -c1-400
fi
