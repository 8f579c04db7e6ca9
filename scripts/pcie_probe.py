"""Host-link ceiling of a box: every rank copies 1.6 GB host->device and 1.2 GB device->host (the bytes of 10^8
queries through gsl_scattered_cuda_eval_e) from / to pinned memory, both directions at once on two streams, and
rank 0 prints the aggregate rate -- what `e2e` in bench.py can reach at most on this box at this N.
    python scripts/pcie_probe.py                                   (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_probe.py"""
import os
import time
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nq = 100_000_000
h_in = torch.empty(2 * nq, dtype=torch.float64).pin_memory()
h_out = torch.empty(nq * 12 // 8, dtype=torch.float64).pin_memory()
d_in = torch.empty_like(h_in, device="cuda")
d_out = torch.empty_like(h_out, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def step():
    with torch.cuda.stream(s_in):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s_out):
        h_out.copy_(d_out, non_blocking=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    step()
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    sec = float(dt.item()) / reps
    print("N=%d: %.1f GB/s host->device + %.1f GB/s device->host per box = %.3e queries/s of 28 B (16 in + 12 out)"
          % (world, world * 1.6 / sec, world * 1.2 / sec, world * nq / sec))
if world > 1:
    dist.destroy_process_group()
