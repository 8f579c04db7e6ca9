#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_devices or golden or slab" > gpurun_out/g2_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/g2_pytest.log
