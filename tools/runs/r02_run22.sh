#!/bin/sh
mkdir -p gpurun_out
python tools/run_jidac_multi.py --gpus 2 > gpurun_out/r02_jidac_multi2.jsonl 2> gpurun_out/r02_jidac_multi2.err; cut -c1-800 gpurun_out/r02_jidac_multi2.jsonl; tail -3 gpurun_out/r02_jidac_multi2.err
