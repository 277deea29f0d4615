"""Golden fixture for the student BEV encoder from the UNMODIFIED reference classes:

    mmdet3d/models/bricks/res_block.py    BasicBlock
    mmdet3d/models/backbones/resnet.py    ResNetForBEVDet
    mmdet3d/models/necks/lss_fpn.py       FPN_LSS

imported from /root/reference with the stubs of tools/ref_import.py (mmcv.cnn build_* -> torch.nn, registries ->
no-ops), built with the shipped config (configs/.../...bevdepth4d_r50.py:122-126) under torch.manual_seed(0), run
in TRAINING mode, fp32, on the CPU: state_dict keys + shapes + per-tensor checksums (our mirrors must create the same
tensors in the same order from the same seed), one input, the output, and the gradients of every parameter
(small tensors in full, large ones as checksums) and of the input. Writes tests/golden/bev_encoder.npz.
Only runs in the build container (/root/reference is not on the GPU box).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402


def load_classes():
    ref_import.install_stubs()
    cnn = sys.modules["mmcv.cnn"]
    cnn.build_plugin_layer = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("plugins"))

    def build_norm_layer(cfg, num_features, postfix=""):
        # mmcv 1.6.0 cnn/bricks/norm.py:build_norm_layer: name = abbreviation ('bn' for BN) + str(postfix); BN = nn.BatchNorm2d
        cfg = dict(cfg)
        assert cfg.pop("type") == "BN"
        cfg.pop("requires_grad", None)
        return "bn" + str(postfix), torch.nn.BatchNorm2d(num_features, **cfg)

    cnn.build_norm_layer = build_norm_layer

    class BaseModule(torch.nn.Module):       # mmcv.runner.BaseModule: nn.Module + init_cfg bookkeeping
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    sys.modules["mmcv.runner"].BaseModule = BaseModule
    mmdet = ref_import._pkg("mmdet")
    mm = ref_import._pkg("mmdet.models")
    mm.BACKBONES = ref_import._Registry("backbone")
    mm.NECKS = ref_import._Registry("neck")
    mmdet.models = mm
    ref_import._pkg("mmdet3d.models.bricks", os.path.join(ref_import.REF_ROOT, "mmdet3d/models/bricks"))
    rb = ref_import.load_ref_module("mmdet3d.models.bricks.res_block", "mmdet3d/models/bricks/res_block.py")
    bricks = sys.modules["mmdet3d.models.bricks"]
    bricks.BasicBlock, bricks.Bottleneck = rb.BasicBlock, rb.Bottleneck
    ref_import._pkg("mmdet3d.models.backbones", os.path.join(ref_import.REF_ROOT, "mmdet3d/models/backbones"))
    rn = ref_import.load_ref_module("mmdet3d.models.backbones.resnet", "mmdet3d/models/backbones/resnet.py")
    fpn = ref_import.load_ref_module("mmdet3d.models.necks.lss_fpn", "mmdet3d/models/necks/lss_fpn.py")
    return rn.ResNetForBEVDet, fpn.FPN_LSS


def checksum(t):
    t = t.detach().double()
    return np.array([float(t.sum()), float(t.abs().sum()), float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64)
                                                                   .reshape(t.shape) % 7).sum())])


def main():
    ResNetForBEVDet, FPN_LSS = load_classes()
    torch.manual_seed(0)
    backbone = ResNetForBEVDet(numC_input=128, num_channels=[128, 256, 512])
    neck = FPN_LSS(in_channels=640, out_channels=256)
    backbone.train(), neck.train()
    init = {}
    for prefix, mod in (("backbone", backbone), ("neck", neck)):
        for k, v in mod.state_dict().items():
            init["sum/" + prefix + "." + k] = checksum(v.float()) if v.dtype != torch.long else np.array([float(v)])
    gen = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(1, 128, 32, 32, generator=gen)).requires_grad_(True)   # the test re-draws x / g from seed 5
    y = neck(backbone(x))
    g = torch.randn(y.shape, generator=gen) / y.numel() ** 0.5
    loss = (y * g).sum()
    loss.backward()
    out = {"x_sum": checksum(x), "g_sum": checksum(g), "y": y.detach().numpy(), "loss": np.array(float(loss.detach())),
           "x_grad": x.grad.numpy()}
    keys, shapes = [], []
    for prefix, mod in (("backbone", backbone), ("neck", neck)):
        for k, v in mod.state_dict().items():
            keys.append(prefix + "." + k)
            shapes.append(list(v.shape))
            if k.endswith("running_mean") or k.endswith("running_var"):
                out["after/" + prefix + "." + k] = v.numpy().copy()          # after ONE training forward
        for k, p in mod.named_parameters():
            name = prefix + "." + k
            if p.numel() <= 4096:
                out["grad/" + name] = p.grad.numpy()
            out["gradsum/" + name] = checksum(p.grad)
            out["gradmax/" + name] = np.array(float(p.grad.abs().max()))
    out.update(init)
    out["keys"] = np.array(keys)
    out["shapes"] = np.array([",".join(map(str, s)) for s in shapes])
    path = os.path.join(ROOT, "tests", "golden", "bev_encoder.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d state_dict entries, %.1f MB)" % (path, len(keys), os.path.getsize(path) / 1e6))


def adaptation_golden():
    """'upsample_3layer' adaptation (bevdet_distill.py:275-301) from the UNMODIFIED ThreeLayer class (:99-130), cut out of
    bevdet_distill.py with `ast` (the module itself needs mmcv / mmdet / cv2): nn.Sequential(nn.Upsample(x4, bilinear,
    align_corners), ThreeLayer(128 -> 128, kernel_size 1, stride 1)) in training mode on the CPU, and an Mlp (:48-68).
    Writes tests/golden/adaptation_3layer.npz."""
    import ast
    from functools import partial
    from torch import nn
    from torch.nn.modules.utils import _pair
    path = os.path.join(ref_import.REF_ROOT, "mmdet3d/models/detectors/bevdet_distill.py")
    tree = ast.parse(open(path).read())
    ns = dict(nn=nn, torch=torch, partial=partial, _pair=_pair, build_norm_layer=None)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("Mlp", "TwoLayer", "ThreeLayer"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    torch.manual_seed(0)
    net = nn.Sequential(nn.Upsample(scale_factor=4, mode="bilinear", align_corners=True),
                        ns["ThreeLayer"](in_features=128, out_features=128, kernel_size=1, stride=1)).train()
    mlp = ns["Mlp"](in_features=128, out_features=128)
    gen = torch.Generator().manual_seed(9)
    x = torch.relu(torch.randn(2, 128, 5, 5, generator=gen)).requires_grad_(True)
    y = net(x)
    g = torch.randn(y.shape, generator=torch.Generator().manual_seed(10))      # the test re-draws it from the seed
    y.backward(g)
    xm = torch.relu(torch.randn(2, 128, 8, 8, generator=gen))
    out = {"x": x.detach().numpy(), "y": y.detach().numpy(), "x_grad": x.grad.numpy(),
           "keys": np.array(list(net.state_dict().keys())), "mlp_keys": np.array(list(mlp.state_dict().keys())),
           "mlp_x": xm.numpy(), "mlp_y": mlp(xm).detach().numpy()}
    for k, v in net.state_dict().items():
        out["sd/" + k] = v.numpy()          # after the training forward: running stats included
    for k, v in mlp.state_dict().items():
        out["mlp_sd/" + k] = v.numpy()
    for k, p in net.named_parameters():
        out["grad/" + k] = p.grad.numpy()
    p = os.path.join(ROOT, "tests", "golden", "adaptation_3layer.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, "%.2f MB" % (os.path.getsize(p) / 1e6))


def bottleneck_golden():
    """ResNetForBEVDet(block_type='BottleNeck') (resnet.py:26-35) with the UNMODIFIED Bottleneck class
    (bricks/res_block.py:102-311): one stage of two blocks, 128 -> 512 channels, stride 2, training mode, fp32 CPU.
    Writes tests/golden/bev_encoder_bottleneck.npz (checksums of the seeded state_dict, output, gradients)."""
    ResNetForBEVDet, _ = load_classes()
    torch.manual_seed(3)
    net = ResNetForBEVDet(numC_input=128, num_layer=[2], num_channels=[512], stride=[2], block_type="BottleNeck").train()
    out = {}
    keys, shapes = [], []
    for k, v in net.state_dict().items():
        keys.append(k)
        shapes.append(",".join(map(str, v.shape)))
        out["sum/" + k] = checksum(v.float()) if v.dtype != torch.long else np.array([float(v)])
    gen = torch.Generator().manual_seed(6)
    x = torch.relu(torch.randn(2, 128, 16, 16, generator=gen)).requires_grad_(True)
    y = net(x)[0]
    g = torch.randn(y.shape, generator=gen) / y.numel() ** 0.5
    loss = (y * g).sum()
    loss.backward()
    out.update(x_sum=checksum(x), g_sum=checksum(g), y=y.detach().numpy(), loss=np.array(float(loss.detach())),
               x_grad=x.grad.numpy(), keys=np.array(keys), shapes=np.array(shapes))
    for k, p in net.named_parameters():
        if p.numel() <= 4096:
            out["grad/" + k] = p.grad.numpy()
        out["gradmax/" + k] = np.array(float(p.grad.abs().max()))
    for k, v in net.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["after/" + k] = v.numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "bev_encoder_bottleneck.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d state_dict entries, %.2f MB)" % (path, len(keys), os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    if "--adaptation" in sys.argv:
        adaptation_golden()
    elif "--bottleneck" in sys.argv:
        bottleneck_golden()
    else:
        main()
