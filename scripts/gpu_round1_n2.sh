#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "two_devices or config1_golden" > gpurun_out/pytest_n2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_samebox.json 2> gpurun_out/bench_n1_samebox.err; echo "bench n1 exit $?"; cat gpurun_out/bench_n1_samebox.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --nodes 100000 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $?"; cat gpurun_out/bench_ref_n2.json
