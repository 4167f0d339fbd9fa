#!/bin/sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -10 > gpurun_out/r02_pytest_final.txt; cat gpurun_out/r02_pytest_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
