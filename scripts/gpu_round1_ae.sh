#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/raster_sweep.py > gpurun_out/ae_raster_sweep.jsonl 2> gpurun_out/ae.err; echo "exit $?"; cat gpurun_out/ae_raster_sweep.jsonl; tail -3 gpurun_out/ae.err
