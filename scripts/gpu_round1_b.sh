#!/bin/bash
# ncu: launch list of the bench command + one full capture of k_locate
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k million > gpurun_out/pytest_million.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_million.log
CMD="python bench.py --queries 100000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-queries 1000000"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
CMD2="python scripts/profile_locate.py 1e6 5e7 1 3"
$CMD2 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -o gpurun_out/prof_locate_r1a $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
