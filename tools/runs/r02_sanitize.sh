#!/bin/sh
# compute-sanitizer over every kernel family (small inputs), logs to gpurun_out/ (copied to profiles/ afterwards)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 6000 > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py 2500 > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck.txt
