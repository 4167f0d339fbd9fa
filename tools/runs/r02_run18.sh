#!/bin/sh
mkdir -p gpurun_out
python tools/run_multi.py --gpus 2 --level 3 --blocks 2048 > gpurun_out/r02_multi2_2048.jsonl 2> gpurun_out/r02_multi2_2048.err; cut -c1-900 gpurun_out/r02_multi2_2048.jsonl; tail -2 gpurun_out/r02_multi2_2048.err
