#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/stress_parity.py 1e6 2e7 > gpurun_out/af_stress_1e6.json 2> gpurun_out/af.err; echo "exit $?"; cat gpurun_out/af_stress_1e6.json; tail -3 gpurun_out/af.err
