#!/bin/bash
# 8-GPU weak-scaling check of both bench arms (the driver runs its own N=1,2,4,8 at round end)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n8_gpus.txt
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/n8_bench_n$N.json 2> gpurun_out/n8_bench_n$N.err; echo "bench n$N exit $?"; wc -l gpurun_out/n8_bench_n$N.json; cut -c1-300 gpurun_out/n8_bench_n$N.json; tail -3 gpurun_out/n8_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/n8_bench_ref_n$N.json 2> gpurun_out/n8_bench_ref_n$N.err; echo "ref n$N exit $?"; wc -l gpurun_out/n8_bench_ref_n$N.json; cut -c1-200 gpurun_out/n8_bench_ref_n$N.json
