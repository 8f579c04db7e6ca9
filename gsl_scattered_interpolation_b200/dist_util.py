"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous only.

The path shards by queries: the triangulation slab is replicated once (broadcast from rank 0
over NCCL/NVLink), every rank evaluates its own contiguous range of the global query stream,
and there is NO collective on the data path.  The only collectives are the slab broadcast at
init and the max-over-ranks of the timed region.
"""
import torch
import torch.distributed as dist


def shard_range(rank, world, per_rank):
    """weak scaling: rank r owns queries k in [r*per_rank, (r+1)*per_rank) of the global stream"""
    return rank * per_rank, (rank + 1) * per_rank


def split_range(rank, world, total):
    """strong scaling: `total` queries split into `world` contiguous slices"""
    return total * rank // world, total * (rank + 1) // world


def broadcast_blob(blob, src=0):
    """Broadcast a 1-D uint8 tensor whose size only `src` knows; returns the received tensor
    (on the same device as `blob`)."""
    n = torch.tensor([blob.numel() if dist.get_rank() == src else 0], dtype=torch.int64,
                     device=blob.device)
    dist.broadcast(n, src)
    if dist.get_rank() != src:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=blob.device)
    dist.broadcast(blob, src)
    return blob


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_over_ranks(value, device):
    """every rank's value, in rank order (diagnostics: which GPU set the max)"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [float(x.item()) for x in out]
    return [float(t.item())]
