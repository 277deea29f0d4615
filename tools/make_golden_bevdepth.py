"""Golden vectors for shift_feature / get_depth_loss, produced by EXECUTING the unmodified method
bodies of mmdet3d/models/detectors/bevdet.py (BEVDetSequentialES.shift_feature :267-321,
BEVDepth_Base.get_depth_loss :397-417), cut out with ast (tools/ref_import.py).

    python tools/make_golden_bevdepth.py   ->  tests/golden/bevdepth_aux.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class _VT(object):
    pass


def main():
    sf, _ = ref_import.load_fgd_methods(names=("shift_feature",), cls_name="BEVDetSequentialES",
                                        relpath="mmdet3d/models/detectors/bevdet.py")
    dl, _ = ref_import.load_fgd_methods(names=("get_depth_loss",), cls_name="BEVDepth_Base",
                                        relpath="mmdet3d/models/detectors/bevdet.py")

    class Fake(object):
        shift_feature = sf["shift_feature"]
        get_depth_loss = dl["get_depth_loss"]
    me = Fake()
    vt = _VT()
    vt.dx = torch.tensor([0.8, 0.8, 20.0])
    vt.bx = torch.tensor([-25.2, -25.2, 0.0])
    vt.D = 19
    vt.grid_config = dict(dbound=[1.0, 20.0, 1.0])
    vt.loss_depth_weight = 3.0
    me.img_view_transformer = vt
    me.interpolation_mode = "bilinear"
    g = torch.Generator().manual_seed(2)
    n, v, c, h, w = 3, 6, 3, 40, 40
    feat = torch.randn(n, c, h, w, generator=g, requires_grad=True)

    def pose(yaw, t):
        r = torch.tensor([[np.cos(yaw), -np.sin(yaw), 0.0], [np.sin(yaw), np.cos(yaw), 0.0], [0, 0, 1.0]],
                         dtype=torch.float32)
        return r, torch.tensor(t, dtype=torch.float32)
    rots0, trans0, rots1, trans1 = [], [], [], []
    for i in range(n):
        cams = [pose(0.3 * k, [0.5 * k, -0.2 * k, 1.5]) for k in range(v)]
        ego_r, ego_t = pose(0.05 * (i + 1), [1.7 * (i + 1), -0.6 * i, 0.0])      # ego motion between the frames
        rots0.append(torch.stack([r for r, _ in cams]))
        trans0.append(torch.stack([t for _, t in cams]))
        rots1.append(torch.stack([ego_r @ r for r, _ in cams]))
        trans1.append(torch.stack([ego_r @ t + ego_t for _, t in cams]))
    rots = [torch.stack(rots0), torch.stack(rots1)]
    trans = [torch.stack(trans0), torch.stack(trans1)]
    out = me.shift_feature(feat, trans, rots)
    wgt = torch.randn(out.shape, generator=g)
    (out * wgt).sum().backward()
    res = dict(sf_in=feat.detach().numpy(), sf_rots0=rots[0].numpy(), sf_rots1=rots[1].numpy(),
               sf_trans0=trans[0].numpy(), sf_trans1=trans[1].numpy(), sf_dx=vt.dx.numpy(), sf_bx=vt.bx.numpy(),
               sf_out=out.detach().numpy(), sf_w=wgt.numpy(), sf_grad=feat.grad.numpy())
    print("shift_feature: out", tuple(out.shape), "nonzero fraction", float((out != 0).float().mean()))
    B, N, H, W = 1, 4, 16, 44
    gt = torch.zeros(B, N, H, W)
    m = torch.rand(B, N, H, W, generator=g) < 0.08
    gt[m] = torch.rand(int(m.sum()), generator=g) * 18.9 + 1.0          # strictly below the last bin edge
    gt[0, 0, 0, :4] = torch.tensor([1.0, 1.999, 0.2, 19.99])              # bin edges / below the first bin
    logits = (torch.randn(B * N, vt.D, H, W, generator=g) * 2).requires_grad_(True)
    loss = me.get_depth_loss(gt, logits)
    loss.backward()
    res.update(dl_gt=gt.numpy(), dl_logits=logits.detach().numpy(), dl_loss=np.float64(float(loss)),
               dl_grad=logits.grad.numpy(), dl_D=np.int64(vt.D), dl_dbound=np.array(vt.grid_config["dbound"]),
               dl_weight=np.float64(vt.loss_depth_weight))
    print("get_depth_loss:", float(loss), "positives", int(m.sum()))
    path = os.path.join(GOLDEN, "bevdepth_aux.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
