"""GPU microbench of the bev_pool plan/gather kernels (CUDA events, L2 flushed).

    python tools/microbench_bev_pool.py [--frames 16] [--bev 128] [--C 64] [--dstep 1.0]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200 import synthetic  # noqa: E402
from oracle import lss_oracle  # noqa: E402  (input geometry only)


def time_ms(fn, iters, flush):
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--bev", type=int, default=128)
    ap.add_argument("--C", type=int, default=64)
    ap.add_argument("--dstep", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--rpi", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    grid = synthetic.grid_config(a.bev, a.dstep)
    dx, bx, nx = lss_oracle.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    frustum = lss_oracle.create_frustum(synthetic.NUSC_INPUT_SIZE, 16, grid["dbound"])
    calib = synthetic.make_calibration(a.frames, 6, seed=a.seed)
    geom = torch.from_numpy(lss_oracle.get_geometry(frustum, *calib)).to(dev)
    n = geom.numel() // 3
    x = torch.rand(n, a.C, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    plan = dbev.bev_plan_from_geom(geom, a.frames, bx, dx, nx, rows_per_item=a.rpi)
    kept = plan.num_kept()
    out = dbev.bev_pool_gather(x, plan)
    og = torch.rand_like(out)
    xr = x.clone().requires_grad_(True)
    for _ in range(3):
        dbev.bev_plan_from_geom(geom, a.frames, bx, dx, nx)
        dbev.bev_pool_gather(x, plan)
    t_plan = time_ms(lambda: dbev.bev_plan_from_geom(geom, a.frames, bx, dx, nx), a.iters, flush)
    t_fwd = time_ms(lambda: dbev.bev_pool_gather(x, plan), a.iters, flush)
    # back-to-back launches without a flush (what bench.py's roofline leg times)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(30):
        dbev.bev_pool_gather(x, plan)
    e1.record()
    torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 30

    def bwd():
        o = dbev.bev_pool_gather(xr, plan)
        o.backward(og)
        xr.grad = None
    t_fb = time_ms(bwd, a.iters, flush)
    fwd_bytes = kept * a.C * 4 + kept * 4 + out.numel() * 4
    bwd_bytes = out.numel() * 4 + n * 4 + n * a.C * 4
    rep = dict(rpi=a.rpi, n_items=int(plan.n_items.item()), frames=a.frames, bev=a.bev, C=a.C, dstep=a.dstep, n_points=n, kept=kept,
               plan_ms_med=t_plan[0], plan_ms_min=t_plan[1], fwd_ms_med=t_fwd[0], fwd_ms_min=t_fwd[1],
               fwd_GBps_med=fwd_bytes / t_fwd[0] / 1e6, fwd_GBps_best=fwd_bytes / t_fwd[1] / 1e6,
               fwdbwd_ms_med=t_fb[0], bwd_ms_est=t_fb[0] - t_fwd[0],
               bwd_GBps_est=bwd_bytes / max(t_fb[0] - t_fwd[0], 1e-6) / 1e6,
               fwd_alg_bytes=fwd_bytes, fwd_ms_b2b=b2b, fwd_GBps_b2b=fwd_bytes / b2b / 1e6)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
