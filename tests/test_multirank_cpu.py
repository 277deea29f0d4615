"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (rank-seeded batches, max-over-ranks
timing, whole-job throughput) and the fact that the hot path shards by sample with no collective:
per-rank oracle results of a 2-sample batch equal the single-process results of the same samples."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import voxel_oracle
    synthetic = bench._load_synthetic()
    # rank-local "step time": rank 1 is slower -> the job is as fast as its slowest rank
    ms = bench.aggregate_step_time(10.0 + 5.0 * rank, world, dist)
    value = bench.samples_per_sec(ms, 1, world)
    # shard-by-sample: rank r voxelizes sample r of the global batch
    cloud = synthetic.make_lidar(world, 5000, seed=3)[rank]
    coors = voxel_oracle.dynamic_voxelize(cloud, bench.PILLAR_VS, bench.PILLAR_RANGE)
    gathered = [None] * world
    dist.all_gather_object(gathered, int(coors.sum()))
    if rank == 0:
        out.put((ms, value, gathered, bench.rank_seed(0), bench.rank_seed(1)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, value, gathered, s0, s1 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 15.0                                   # MAX over ranks, not mean
    assert abs(value - 8 * 2 / 15e-3) < 1e-6            # whole-job samples/s (weak scaling)
    assert s0 != s1                                     # ranks draw different batches
    sys.path.insert(0, ROOT)
    import bench
    from oracle import voxel_oracle
    synthetic = bench._load_synthetic()
    clouds = synthetic.make_lidar(2, 5000, seed=3)
    expect = [int(voxel_oracle.dynamic_voxelize(c, bench.PILLAR_VS, bench.PILLAR_RANGE).sum()) for c in clouds]
    assert gathered == expect                           # per-rank shards == single-process batch
