#!/bin/sh
# Round-end measurement pass on one B200 (round 2): tests, smoke, bench (both arms, every BASELINE configuration in
# per_level), the ncu launch list of the bench command (one metric, one pass), the DRAM traffic of the codec kernels on
# the bench configuration and full captures of the two codec kernels (64 KiB blocks: short enough for ncu's replays).
# Outputs under gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_pytest_final.txt; cat gpurun_out/r02_pytest_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err; head -c 700 gpurun_out/r02_bench_final.json; echo
python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2>/dev/null; head -c 400 gpurun_out/r02_bench_ref.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-per-level > gpurun_out/r02_bench_under_ncu.json 2>/dev/null
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_(en|de)code' --csv --log-file gpurun_out/r02_traffic_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-per-level > gpurun_out/r02_traffic_bench.json 2> gpurun_out/r02_traffic.err
B="python bench.py --blocks 1024 --block-kib 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-level"
ncu --set full --clock-control none --import-source on -k 'regex:k_decode' -s 1 -c 1 -o gpurun_out/r02_final_dec $B > gpurun_out/r02_ncu_final_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_encode' -s 1 -c 1 -o gpurun_out/r02_final_enc $B > gpurun_out/r02_ncu_final_enc.log 2>&1
ls -la gpurun_out/r02_final_*.ncu-rep gpurun_out/r02_bench_launches.csv
