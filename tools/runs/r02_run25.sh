#!/bin/sh
# last binary of the round: the multi-device handle tests (host threads) and smoke
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
