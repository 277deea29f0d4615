"""GPU microbench of the fused lift+splat kernels at the bench shape (16 sample-frames)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200 import synthetic  # noqa: E402


def main(frames=16, iters=10):
    dev = torch.device("cuda:0")
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=8).to(dev)
    calib = [torch.from_numpy(a).to(dev) for a in synthetic.make_calibration(frames, 6, seed=0)]
    geom = vt.get_geometry(*calib)
    depth = torch.randn(frames * 6, 59, 16, 44, device=dev).softmax(1).requires_grad_(True)
    feat = torch.randn(frames * 6, 64, 16, 44, device=dev).requires_grad_(True)
    og = torch.rand(frames, 64, 128, 128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for name, fn in (("plan", lambda: vt.make_plan(geom, frames)),):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name + "_ms"] = sorted(ts)[len(ts) // 2]
    plan = vt.make_plan(geom, frames)
    tf, tb = [], []
    for i in range(iters + 3):
        flush.zero_()
        a, b, c = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        out = dbev.lift_splat(depth, feat, plan)
        b.record()
        out.backward(og)
        c.record()
        torch.cuda.synchronize()
        depth.grad = None; feat.grad = None
        if i >= 3:
            tf.append(a.elapsed_time(b)); tb.append(b.elapsed_time(c))
    res["fwd_ms"] = sorted(tf)[len(tf) // 2]
    res["bwd_ms"] = sorted(tb)[len(tb) // 2]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
