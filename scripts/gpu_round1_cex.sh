#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_c_example.py -q > gpurun_out/cex_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/cex_pytest.log
./oracle/_ref/glue_demo 20000 7 512
