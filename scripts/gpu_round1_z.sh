#!/bin/bash
# 2-GPU check of both bench arms under torchrun + smoke + GPU tests at HEAD
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/z_gpus.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/z_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_z.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_z.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/z_bench_n2.json 2> gpurun_out/z_bench_n2.err; echo "bench n2 exit $?"; wc -l gpurun_out/z_bench_n2.json; cut -c1-400 gpurun_out/z_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/z_bench_ref_n2.json 2> gpurun_out/z_bench_ref_n2.err; echo "ref n2 exit $?"; wc -l gpurun_out/z_bench_ref_n2.json; cut -c1-300 gpurun_out/z_bench_ref_n2.json
