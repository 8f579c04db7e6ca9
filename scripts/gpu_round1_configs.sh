#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py 2 3 4 > gpurun_out/cfg_final.jsonl 2> gpurun_out/cfg_final.err; echo "exit $?"; cut -c1-420 gpurun_out/cfg_final.jsonl; tail -2 gpurun_out/cfg_final.err
