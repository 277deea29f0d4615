"""Import UNMODIFIED reference Python files from /root/reference with stub deps.

Only for the build container (tools/make_golden.py): /root/reference does not
exist on the GPU box, so nothing in tests/, smoke() or bench.py imports this.

The reference needs mmcv / mmdet / torch_scatter, none of which are installed
(SURVEY.md "Facts established by probing"). The stubs below provide exactly
the names the imported files touch at import time; none of them carries
arithmetic except ``torch_scatter.scatter_sum`` which is index_add (the
published semantics of torch_scatter's scatter_sum; parity unpinned for that
third-party call).
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = "/root/reference"


class _Registry(object):
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def build(self, cfg, **kw):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop("type")](**cfg, **kw)


def _pkg(name, path=None):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
    return m


def install_stubs():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_dbev_stub", False):
        return
    mmcv = _pkg("mmcv")
    mmcv._dbev_stub = True
    runner = _pkg("mmcv.runner")
    runner.BaseModule = nn.Module
    runner.force_fp32 = lambda *a, **k: (lambda f: f)
    runner.auto_fp16 = lambda *a, **k: (lambda f: f)
    mmcv.runner = runner
    cnn = _pkg("mmcv.cnn")

    def build_conv_layer(cfg, *args, **kwargs):
        if cfg is not None and cfg.get("type") not in (None, "Conv2d"):
            raise NotImplementedError("stub build_conv_layer: %s" % cfg)
        return nn.Conv2d(*args, **kwargs)

    def build_norm_layer(cfg, num_features, postfix=""):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        cfg.pop("requires_grad", None)
        cls = {"BN": nn.BatchNorm2d, "BN2d": nn.BatchNorm2d, "BN1d": nn.BatchNorm1d}[typ]
        return "bn", cls(num_features, **cfg)

    def build_activation_layer(cfg):
        assert cfg["type"] == "ReLU"
        return nn.ReLU(inplace=cfg.get("inplace", False))

    cnn.build_activation_layer = build_activation_layer

    cnn.build_conv_layer = build_conv_layer
    cnn.build_norm_layer = build_norm_layer
    cnn.CONV_LAYERS = _Registry("conv layer")
    mmcv.cnn = cnn

    ts = _pkg("torch_scatter")

    def scatter_sum(src, index, dim=0, out=None, dim_size=None):
        assert dim == 0
        size = int(index.max()) + 1 if dim_size is None else dim_size
        res = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        return res.index_add_(0, index, src)

    ts.scatter_sum = scatter_sum

    # package skeleton so that relative imports inside the reference files resolve
    _pkg("mmdet3d", os.path.join(REF_ROOT, "mmdet3d"))
    models = _pkg("mmdet3d.models", os.path.join(REF_ROOT, "mmdet3d/models"))
    builder = _pkg("mmdet3d.models.builder")
    builder.NECKS = _Registry("neck")
    builder.BACKBONES = _Registry("backbone")
    builder.build_backbone = lambda cfg: builder.BACKBONES.build(cfg)
    models.builder = builder
    _pkg("mmdet3d.models.necks", os.path.join(REF_ROOT, "mmdet3d/models/necks"))
    _pkg("mmdet3d.ops", os.path.join(REF_ROOT, "mmdet3d/ops"))
    bp = _pkg("mmdet3d.ops.bev_pool", os.path.join(REF_ROOT, "mmdet3d/ops/bev_pool"))
    ext = _pkg("mmdet3d.ops.bev_pool.bev_pool_ext")  # CUDA-only in the reference; never called here
    bp.bev_pool_ext = ext


def load_ref_module(qualname, relpath):
    """exec the reference file at REF_ROOT/relpath as module `qualname`."""
    install_stubs()
    if qualname in sys.modules and getattr(sys.modules[qualname], "__file__", None):
        return sys.modules[qualname]
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(qualname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[qualname] = mod
    spec.loader.exec_module(mod)
    return mod


class _CpuDynamicScatter(nn.Module):
    """CPU stand-in for mmdet3d.ops.DynamicScatter (CUDA-only in the reference,
    voxelization.h:118): the host sequence of scatter_points_cuda.cu:183-239 executed with
    torch CPU ops (masked_fill, torch.unique(dim=0) = at::unique_dim, scatter amax / index_add),
    wrapped in the reference's own per-sample loop (scatter_points.py:75-100)."""

    def __init__(self, voxel_size, point_cloud_range, average_points):
        super().__init__()
        self.average_points = average_points

    def forward_single(self, points, coors):
        clean = coors.masked_fill(coors.lt(0).any(-1, True), -1)
        oc, inv, cnt = torch.unique(clean, dim=0, sorted=True, return_inverse=True, return_counts=True)
        if oc[0, 0] < 0:
            oc, cnt, inv = oc[1:], cnt[1:], inv - 1
        valid = inv >= 0
        m, c = oc.shape[0], points.shape[1]
        if self.average_points:
            red = torch.zeros(m, c).index_add_(0, inv[valid], points[valid]) / cnt[:, None].float()
        else:
            red = torch.full((m, c), -float("inf")).scatter_reduce(
                0, inv[valid][:, None].expand(-1, c), points[valid], "amax")
        return red, oc

    def forward(self, points, coors):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        batch_size = int(coors[-1, 0] + 1)
        voxels, voxel_coors = [], []
        for i in range(batch_size):
            inds = torch.where(coors[:, 0] == i)
            voxel, voxel_coor = self.forward_single(points[inds], coors[inds][:, 1:])
            voxel_coors.append(nn.functional.pad(voxel_coor, (1, 0), mode="constant", value=i))
            voxels.append(voxel)
        return torch.cat(voxels, dim=0), torch.cat(voxel_coors, dim=0)


def pillar_modules():
    """(DynamicPillarFeatureNet, PointPillarsScatter) classes from the unmodified reference files."""
    install_stubs()
    sys.modules["mmdet3d.ops"].DynamicScatter = _CpuDynamicScatter
    builder = sys.modules["mmdet3d.models.builder"]
    builder.VOXEL_ENCODERS = _Registry("voxel encoder")
    builder.MIDDLE_ENCODERS = _Registry("middle encoder")
    _pkg("mmdet3d.models.voxel_encoders", os.path.join(REF_ROOT, "mmdet3d/models/voxel_encoders"))
    _pkg("mmdet3d.models.middle_encoders", os.path.join(REF_ROOT, "mmdet3d/models/middle_encoders"))
    load_ref_module("mmdet3d.models.voxel_encoders.utils", "mmdet3d/models/voxel_encoders/utils.py")
    pe = load_ref_module("mmdet3d.models.voxel_encoders.pillar_encoder",
                         "mmdet3d/models/voxel_encoders/pillar_encoder.py")
    ps = load_ref_module("mmdet3d.models.middle_encoders.pillar_scatter",
                         "mmdet3d/models/middle_encoders/pillar_scatter.py")
    return pe.DynamicPillarFeatureNet, ps.PointPillarsScatter


def view_transformer_mine():
    return load_ref_module("mmdet3d.models.necks.view_transformer_mine",
                           "mmdet3d/models/necks/view_transformer_mine.py")


def bev_pool_py():
    return load_ref_module("mmdet3d.ops.bev_pool.bev_pool", "mmdet3d/ops/bev_pool/bev_pool.py")


# --------------------------------------------------------------------------
# DistillBEV loss methods: the file mmdet3d/models/detectors/bevdet_distill.py
# cannot be imported (mmcv, mmdet, cv2, sibling detectors ...), so the
# UNMODIFIED source of the methods we need is cut out of the class with `ast`
# and compiled in a namespace holding the real reference helpers
# (box_np_ops, LiDARPoints, clip_sigmoid) plus a stand-in for mmdet's
# build_loss (third-party mmdet==2.24.0: MSELoss / L1Loss / SmoothL1Loss with
# reduction='none' are loss_weight * F.{mse,l1,smooth_l1}_loss(..., 'none');
# parity unpinned for that dependency).
# --------------------------------------------------------------------------

class _AttrDict(dict):
    """mmcv ConfigDict stand-in: dict with attribute access, getattr(default) works."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _build_loss(cfg):
    import torch.nn.functional as F
    cfg = dict(cfg)
    typ = cfg.pop("type")
    red = cfg.pop("reduction", "mean")
    lw = cfg.pop("loss_weight", 1.0)
    fn = {"MSELoss": F.mse_loss, "L1Loss": F.l1_loss, "SmoothL1Loss": F.smooth_l1_loss}[typ]
    return lambda pred, target: lw * fn(pred, target, reduction=red)


def load_fgd_methods(names=("foreground_scale_mask", "add_fp_as_fg", "fgd_distill_loss",
                            "affinity_distill_loss"), cls_name="BEVDetDistill",
                     relpath="mmdet3d/models/detectors/bevdet_distill.py", extra_ns=None):
    """dict name -> python function(self, ...) compiled from the reference source."""
    import ast
    import copy
    import functools
    import numpy as np
    import torch.nn.functional as F
    install_stubs()
    box_np_ops = load_ref_module("ref_box_np_ops", "mmdet3d/core/bbox/box_np_ops.py")
    _pkg("ref_points", os.path.join(REF_ROOT, "mmdet3d/core/points"))
    load_ref_module("ref_points.base_points", "mmdet3d/core/points/base_points.py")
    lidar_points = load_ref_module("ref_points.lidar_points", "mmdet3d/core/points/lidar_points.py")
    clip = load_ref_module("ref_clip_sigmoid", "mmdet3d/models/utils/clip_sigmoid.py")
    path = os.path.join(REF_ROOT, relpath)
    src = open(path).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
    ns = dict(torch=torch, nn=nn, F=F, np=np, deepcopy=copy.deepcopy, partial=functools.partial,
              box_np_ops=box_np_ops, LiDARPoints=lidar_points.LiDARPoints,
              clip_sigmoid=clip.clip_sigmoid, build_loss=_build_loss, os=os)
    if extra_ns:
        ns.update(extra_ns)
    out = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            code = compile(mod, path, "exec")  # keeps reference file:line in tracebacks
            exec(code, ns)
            out[node.name] = ns[node.name]
    missing = set(names) - set(out)
    assert not missing, missing
    return out, _AttrDict
