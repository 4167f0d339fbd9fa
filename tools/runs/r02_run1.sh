#!/bin/sh
# round 2, first GPU pass: tests, bench with per_level, DRAM traffic of the codec kernels on the bench configuration
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -c 1500 gpurun_out/r02_bench1.err; head -c 3000 gpurun_out/r02_bench1.json
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_(en|de)code' --csv --log-file gpurun_out/r02_traffic_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-per-level > gpurun_out/r02_traffic_bench.json 2> gpurun_out/r02_traffic.err
tail -5 gpurun_out/r02_traffic_launches.csv
