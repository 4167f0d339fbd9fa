#!/bin/sh
# last pass of round 2 on the final binary: the whole GPU test suite, smoke, the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -12 > gpurun_out/r02_pytest_final.txt; cat gpurun_out/r02_pytest_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err; head -c 400 gpurun_out/r02_bench_final.json; echo
