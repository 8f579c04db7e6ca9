"""The N > 1 path on CPU: world_size-2 gloo processes run the same sharding / replication /
timing-reduction plumbing as bench.py (gsl_scattered_interpolation_b200/dist_util.py), with
the oracle port standing in for the GPU kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gsl_scattered_interpolation_b200 as g
    import refimpl
    from gsl_scattered_interpolation_b200 import dist_util, synth

    n, per_rank = 2000, 3000
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, 1)
    # rank 0 builds the triangulation on the host, the others receive it as one blob
    if rank == 0:
        w = g.ScatteredCuda(n, 1)
        w.build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
        t = w.get_tree(w.n_simplexes(), n)
        w.free()
        keys = ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")
        parts = [np.ascontiguousarray(t[k]).view(np.uint8) for k in keys]
        sizes = np.array([p.size for p in parts], dtype=np.int64)
        blob = torch.from_numpy(np.concatenate([sizes.view(np.uint8)] + parts))
    else:
        blob = torch.empty(0, dtype=torch.uint8)
    blob = dist_util.broadcast_blob(blob, 0).numpy()
    sizes = blob[:56].view(np.int64)
    offs = 56 + np.concatenate([[0], np.cumsum(sizes)])
    dt = (np.int32, np.int32, np.int32, np.uint64, np.float64, np.float64, np.float64)
    keys = ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")
    tree = {k: blob[offs[i]:offs[i + 1]].view(dt[i]).copy() for i, k in enumerate(keys)}
    # every rank evaluates its own contiguous range of the global query stream
    k0, k1 = dist_util.shard_range(rank, world, per_rank)
    xq, yq = synth.uniform_queries(k1 - k0, first=k0)
    port_tree = refimpl.PortTree(tree, xa, ya)
    leaf, val, _ = port_tree.eval(xq, yq, za)
    np.save(os.path.join(out_dir, f"leaf{rank}.npy"), leaf)
    np.save(os.path.join(out_dir, f"val{rank}.npy"), val)
    # timing reduction: max over ranks
    m = dist_util.max_over_ranks(10.0 + rank, torch.device("cpu"))
    assert m == 10.0 + world - 1
    assert dist_util.gather_over_ranks(10.0 + rank, torch.device("cpu")) == [10.0 + r for r in range(world)]
    s0, s1 = dist_util.split_range(rank, world, 1001)
    tot = torch.tensor([s1 - s0]); dist.all_reduce(tot)
    assert int(tot.item()) == 1001
    # strong scaling (bench.py's scaling_strong leg): a FIXED stream of 5001 queries cut into `world` slices
    s0, s1 = dist_util.split_range(rank, world, 5001)
    xq, yq = synth.uniform_queries(s1 - s0, first=s0)
    leaf, val, _ = port_tree.eval(xq, yq, za)
    np.save(os.path.join(out_dir, f"sleaf{rank}.npy"), leaf)
    np.save(os.path.join(out_dir, f"sval{rank}.npy"), val)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gsl_scattered_interpolation_b200 as g
    import refimpl
    from gsl_scattered_interpolation_b200 import synth
    world, per_rank = 2, 3000
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n = 2000
    xa, ya = synth.uniform_nodes(n)
    za = synth.values(n, 1)
    w = g.ScatteredCuda(n, 1)
    w.build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    tree = w.get_tree(w.n_simplexes(), n)
    w.free()
    xq, yq = synth.uniform_queries(world * per_rank)
    leaf, val, _ = refimpl.PortTree(tree, xa, ya).eval(xq, yq, za)
    got_leaf = np.concatenate([np.load(tmp_path / f"leaf{r}.npy") for r in range(world)])
    got_val = np.concatenate([np.load(tmp_path / f"val{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(got_leaf, leaf)
    assert refimpl.same_bits(got_val, val)
    xq, yq = synth.uniform_queries(5001)
    leaf, val, _ = refimpl.PortTree(tree, xa, ya).eval(xq, yq, za)
    got_leaf = np.concatenate([np.load(tmp_path / f"sleaf{r}.npy") for r in range(world)])
    got_val = np.concatenate([np.load(tmp_path / f"sval{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(got_leaf, leaf) and refimpl.same_bits(got_val, val)
