#!/bin/sh
# SHA-1 kernel A/B, then the whole GPU test suite (with the 16 MiB single-block case) and smoke on the same binary
mkdir -p gpurun_out
python tools/ab_sha1.py > gpurun_out/r02_ab_sha1.jsonl 2> gpurun_out/r02_ab_sha1.err; cat gpurun_out/r02_ab_sha1.jsonl; tail -3 gpurun_out/r02_ab_sha1.err
timeout 320 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 > gpurun_out/r02_pytest_final2.txt; cat gpurun_out/r02_pytest_final2.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
