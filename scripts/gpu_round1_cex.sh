#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_c_example.py tests/test_capi.py -q > gpurun_out/cex_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/cex_pytest.log
