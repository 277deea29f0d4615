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
        return "bn", nn.BatchNorm2d(num_features)

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


def view_transformer_mine():
    return load_ref_module("mmdet3d.models.necks.view_transformer_mine",
                           "mmdet3d/models/necks/view_transformer_mine.py")


def bev_pool_py():
    return load_ref_module("mmdet3d.ops.bev_pool.bev_pool", "mmdet3d/ops/bev_pool/bev_pool.py")
