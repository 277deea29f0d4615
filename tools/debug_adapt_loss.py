"""Fused adaptation + loss (csrc/adapt_loss_tc.cu) against the unfused composition: per-quantity errors and timings."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200 import synthetic  # noqa: E402
from distill_bev_b200.plugin.distill import fgd  # noqa: E402
from distill_bev_b200.plugin.distill.adaptation import Conv1x1Adaptation  # noqa: E402
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, Cs, Ct, H = 8, 256, 384, 128
    rng = np.random.RandomState(0)
    teacher = torch.from_numpy(np.maximum(rng.randn(B, Ct, H, H), 0).astype(np.float32)).to(dev)
    student = torch.from_numpy(np.maximum(rng.randn(B, Cs, H, H), 0).astype(np.float32)).to(dev).contiguous(memory_format=torch.channels_last)
    boxes = [torch.from_numpy(b) for b, _ in synthetic.make_gt_boxes(B, seed=5)]
    tc = dict(grid_size=[1024, 1024, 40], point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], voxel_size=[0.1, 0.1, 0.2])
    p = dict(bench.DISTILL_PARAMS, fp_as_foreground=["none"], fp_weight=0.0)
    torch.manual_seed(1)
    spatial = torch.nn.Conv2d(1, 1, 3, padding=1).to(dev)
    adapt = Conv1x1Adaptation(Cs, Ct).to(dev)

    def fused():
        s = student.clone().requires_grad_(True)
        adapt.zero_grad(), spatial.zero_grad()
        l = fgd.fgd_distill_loss(teacher, s, boxes, p, tc, channel_adaptation=adapt, spatial_adaptation=spatial)
        sum(l.values()).backward()
        return l, [s.grad, adapt.weight.grad.clone(), adapt.bias.grad.clone(), spatial.weight.grad.clone(), spatial.bias.grad.clone()]

    def unfused():
        s = student.clone().requires_grad_(True)
        adapt.zero_grad(), spatial.zero_grad()
        l = fgd.fgd_distill_loss(teacher, adapt(s), boxes, p, tc, spatial_adaptation=spatial)
        sum(l.values()).backward()
        return l, [s.grad, adapt.weight.grad.clone(), adapt.bias.grad.clone(), spatial.weight.grad.clone(), spatial.bias.grad.clone()]

    l1, g1 = fused()
    l2, g2 = unfused()
    for k in l2:
        print(k, float(l1[k]), float(l2[k]))
    for n, a, b in zip(["dx", "dW", "dbias", "dconv_w", "dconv_b"], g1, g2):
        print(n, "max abs err %.3e  max |ref| %.3e  rel %.3e" % (float((a - b).abs().max()), float(b.abs().max()),
                                                                  float((a - b).abs().max() / b.abs().max())))
    for name, fn in (("fused", fused), ("unfused", unfused)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(name, "fwd+bwd incl. masks, eager: %.3f ms" % (a.elapsed_time(b) / 10))


if __name__ == "__main__":
    main()
