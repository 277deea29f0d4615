"""Golden vectors for the sparse LiDAR teacher path, produced by EXECUTING the reference.

Run in the build container only (needs /root/reference and oracle/_ref/ref_sparse_conv_ext.so,
the reference's vendored spconv extension compiled unmodified by oracle/build_oracle.py):

    python tools/make_golden_sparse.py      ->  tests/golden/sparse_small.npz

What runs, unmodified (import stubs only; tools/ref_import.py):
  mmdet3d/ops/spconv/{__init__,conv,functional,modules,ops,pool,structure}.py on top of the
  compiled reference extension (CPU branch), mmdet3d/ops/sparse_block.py,
  mmdet3d/models/middle_encoders/sparse_encoder.py, mmdet3d/models/voxel_encoders/
  dynamic_voxel_encoder.py (+ mmdet3d/core/utils/scatter.py) and HardSimpleVFE from
  voxel_encoders/voxel_encoder.py.
Third party, absent and restated here (parity unpinned at that boundary): mmdet 2.24
`BasicBlock.__init__` (the module layout SparseBasicBlock inherits), mmcv `build_conv_layer` /
`build_norm_layer` (registry lookups).

Weights are NOT stored: both this script and the tests derive them from
oracle.spconv_oracle.fill_params(specs, seed).
"""
import glob
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
from oracle import spconv_oracle as so  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = ref_import.REF_ROOT


def load_ref_ext():
    path = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_sparse_conv_ext*.so"))[0]
    spec = importlib.util.spec_from_file_location("ref_sparse_conv_ext", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install():
    ref_import.install_stubs()
    cnn = sys.modules["mmcv.cnn"]
    registry = cnn.CONV_LAYERS

    def build_conv_layer(cfg, *args, **kwargs):
        cfg = dict(cfg) if cfg is not None else dict(type="Conv2d")
        typ = cfg.pop("type")
        if typ in registry.module_dict:
            return registry.module_dict[typ](*args, **kwargs, **cfg)
        assert typ == "Conv2d", typ
        return nn.Conv2d(*args, **kwargs, **cfg)

    def build_norm_layer(cfg, num_features, postfix=""):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        cfg.pop("requires_grad", None)
        cls = {"BN": nn.BatchNorm2d, "BN2d": nn.BatchNorm2d, "BN1d": nn.BatchNorm1d}[typ]
        return "bn%s" % postfix, cls(num_features, **cfg)

    cnn.build_conv_layer, cnn.build_norm_layer = build_conv_layer, build_norm_layer

    # the spconv python package, unmodified, on the reference's own extension
    pkg = ref_import._pkg("mmdet3d.ops.spconv", os.path.join(REF, "mmdet3d/ops/spconv"))
    pkg.sparse_conv_ext = load_ref_ext()
    sys.modules["mmdet3d.ops.spconv.sparse_conv_ext"] = pkg.sparse_conv_ext
    for name in ("structure", "ops", "functional", "modules", "conv", "pool"):
        m = ref_import.load_ref_module("mmdet3d.ops.spconv." + name,
                                       "mmdet3d/ops/spconv/%s.py" % name)
        setattr(pkg, name, m)
    init_src = open(os.path.join(REF, "mmdet3d/ops/spconv/__init__.py")).read()
    exec(compile(init_src, os.path.join(REF, "mmdet3d/ops/spconv/__init__.py"), "exec"), pkg.__dict__)
    ops_pkg = sys.modules["mmdet3d.ops"]
    ops_pkg.spconv = pkg

    # mmdet 2.24 BasicBlock / Bottleneck constructor layout (third party, restated)
    mmdet = ref_import._pkg("mmdet")
    ref_import._pkg("mmdet.models")
    ref_import._pkg("mmdet.models.backbones")
    resnet = ref_import._pkg("mmdet.models.backbones.resnet")

    class BasicBlock(nn.Module):
        expansion = 1

        def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None,
                     style="pytorch", with_cp=False, conv_cfg=None, norm_cfg=dict(type="BN"),
                     dcn=None, plugins=None, init_cfg=None):
            nn.Module.__init__(self)
            self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
            self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
            self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 3, stride=stride,
                                          padding=dilation, dilation=dilation, bias=False)
            self.add_module(self.norm1_name, norm1)
            self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, padding=1, bias=False)
            self.add_module(self.norm2_name, norm2)
            self.relu = nn.ReLU(inplace=True)
            self.downsample = downsample

        @property
        def norm1(self):
            return getattr(self, self.norm1_name)

        @property
        def norm2(self):
            return getattr(self, self.norm2_name)

    class Bottleneck(BasicBlock):
        expansion = 4

    resnet.BasicBlock, resnet.Bottleneck = BasicBlock, Bottleneck
    sb = ref_import.load_ref_module("mmdet3d.ops.sparse_block", "mmdet3d/ops/sparse_block.py")
    ops_pkg.SparseBasicBlock = sb.SparseBasicBlock
    ops_pkg.make_sparse_convmodule = sb.make_sparse_convmodule
    builder = sys.modules["mmdet3d.models.builder"]
    builder.MIDDLE_ENCODERS = ref_import._Registry("middle encoder")
    builder.VOXEL_ENCODERS = ref_import._Registry("voxel encoder")
    ref_import._pkg("mmdet3d.models.middle_encoders", os.path.join(REF, "mmdet3d/models/middle_encoders"))
    se = ref_import.load_ref_module("mmdet3d.models.middle_encoders.sparse_encoder",
                                    "mmdet3d/models/middle_encoders/sparse_encoder.py")
    # DynamicVoxelEncoder + its scatter_mean
    ref_import._pkg("mmdet3d.core", os.path.join(REF, "mmdet3d/core"))
    ref_import._pkg("mmdet3d.core.utils", os.path.join(REF, "mmdet3d/core/utils"))
    ref_import.load_ref_module("mmdet3d.core.utils.scatter", "mmdet3d/core/utils/scatter.py")
    ref_import._pkg("mmdet3d.models.voxel_encoders", os.path.join(REF, "mmdet3d/models/voxel_encoders"))
    dve = ref_import.load_ref_module("mmdet3d.models.voxel_encoders.dynamic_voxel_encoder",
                                     "mmdet3d/models/voxel_encoders/dynamic_voxel_encoder.py")
    return pkg, se, dve


def random_voxels(rs, batch, shape, n_per_sample, nfeat):
    """Clustered active voxels (so that neighbours exist), unique per sample, (b,z,y,x) int32."""
    Z, Y, X = shape
    coors = []
    for b in range(batch):
        centres = rs.uniform(0, 1, (8, 3)) * np.array([Z, Y, X])
        pts = centres[rs.randint(0, 8, n_per_sample * 2)] + rs.standard_normal((n_per_sample * 2, 3)) * \
            np.array([1.5, 3.0, 3.0])
        c = np.floor(pts).astype(np.int64)
        ok = (c >= 0).all(1) & (c[:, 0] < Z) & (c[:, 1] < Y) & (c[:, 2] < X)
        c = np.unique(c[ok], axis=0)
        c = c[rs.permutation(len(c))[:n_per_sample]]
        coors.append(np.concatenate([np.full((len(c), 1), b), c], 1))
    coors = np.concatenate(coors, 0).astype(np.int32)
    feats = rs.standard_normal((len(coors), nfeat)).astype(np.float32)
    return feats, coors


def set_encoder_params(enc, specs):
    """Copy the seeded parameters of `specs` into the reference SparseEncoder, matching convs and
    norms by traversal (= execution) order."""
    convs = [m for m in enc.modules() if type(m).__name__ in ("SubMConv3d", "SparseConv3d")]
    bns = [m for m in enc.modules() if isinstance(m, nn.BatchNorm1d)]
    flat = list(so._iter_convs(specs))
    assert len(convs) == len(flat) == len(bns), (len(convs), len(flat), len(bns))
    with torch.no_grad():
        for m, bn, c in zip(convs, bns, flat):
            assert tuple(m.weight.shape) == c["weight"].shape, (m.weight.shape, c["weight"].shape)
            assert list(m.kernel_size) == c["ksize"] and m.subm == c["subm"], (m, c["ksize"])
            if not m.subm:
                assert list(m.stride) == c["stride"] and list(m.padding) == c["padding"]
            assert m.indice_key == c["indice_key"]
            m.weight.copy_(torch.from_numpy(c["weight"]))
            bn.weight.copy_(torch.from_numpy(c["bn"]["weight"]))
            bn.bias.copy_(torch.from_numpy(c["bn"]["bias"]))
            bn.running_mean.copy_(torch.from_numpy(c["bn"]["mean"]))
            bn.running_var.copy_(torch.from_numpy(c["bn"]["var"]))
            assert abs(bn.eps - c["bn"]["eps"]) < 1e-12


ENCODERS = {
    # LidarFormer teacher layout (configs/teacher_transformer/lidarformer.py:43-51), small grid
    "lf": dict(in_channels=5, sparse_shape=[41, 48, 48], output_channels=128,
               encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)),
               block_type="basicblock"),
    # SparseEncoder defaults (SECOND layout), keyed rulebook reuse inside a stage
    "sec": dict(in_channels=4, sparse_shape=[41, 32, 32], output_channels=128,
                encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                block_type="conv_module"),
}


def main():
    sp, se, dve = install()
    out = {}
    rs = np.random.RandomState(7)

    # ---- rulebooks straight from the extension (CPU branch) ----
    cases = [("subm3", [9, 16, 16], [3, 3, 3], [1, 1, 1], [1, 1, 1], True),
             ("subm133", [9, 16, 16], [1, 3, 3], [1, 1, 1], [0, 1, 1], True),
             ("s2p1", [9, 16, 16], [3, 3, 3], [2, 2, 2], [1, 1, 1], False),
             ("s2p0", [9, 16, 16], [3, 3, 3], [2, 2, 2], [0, 0, 0], False),
             ("s2p011", [11, 16, 16], [3, 3, 3], [2, 2, 2], [0, 1, 1], False),
             ("down311", [5, 16, 16], [3, 1, 1], [2, 1, 1], [0, 0, 0], False),
             ("s1p1", [9, 16, 16], [3, 3, 3], [1, 1, 1], [1, 1, 1], False)]
    names = []
    for name, shape, k, s, p, subm in cases:
        feats, coors = random_voxels(rs, 2, shape, 150, 8)
        outids, pairs, num = sp.ops.get_indice_pairs(torch.from_numpy(coors), 2, shape, k, s, p, 1, 0, subm)
        w = (rs.standard_normal(tuple(k) + (8, 16)) * 0.2).astype(np.float32)
        y = sp.ops.indice_conv(torch.from_numpy(feats), torch.from_numpy(w), pairs, num,
                               outids.shape[0], False, subm)
        o_out, o_pairs, o_num = so.get_indice_pairs(coors, 2, shape, k, s, p, [1, 1, 1], subm)
        assert np.array_equal(o_out, outids.numpy()), name
        assert np.array_equal(o_num, num.numpy()), name
        assert np.array_equal(o_pairs, pairs.numpy()), name
        oy = so.indice_conv(feats, w, o_pairs, o_num, len(o_out))
        assert np.abs(oy - y.numpy()).max() < 1e-4, name
        names.append(name)
        out["rb_%s_geom" % name] = np.array(shape + k + s + p + [int(subm)], dtype=np.int32)
        out["rb_%s_coors" % name] = coors
        out["rb_%s_feats" % name] = feats
        out["rb_%s_w" % name] = w
        out["rb_%s_outids" % name] = outids.numpy()
        out["rb_%s_pairs" % name] = pairs.numpy()
        out["rb_%s_num" % name] = num.numpy()
        out["rb_%s_y" % name] = y.numpy()
        print("rulebook %-8s n_in %d n_out %d pairs %d: oracle == reference" %
              (name, len(coors), outids.shape[0], int(num.sum())))
    out["rb_names"] = np.array(names)

    # ---- whole SparseEncoder, eval mode ----
    for tag, cfg in ENCODERS.items():
        torch.manual_seed(0)
        enc = se.SparseEncoder(**cfg).eval()
        specs = so.fill_params(so.encoder_layer_specs(
            cfg["in_channels"], 16, cfg["output_channels"], cfg["encoder_channels"],
            cfg["encoder_paddings"], cfg["block_type"]), seed=11)
        set_encoder_params(enc, specs)
        feats, coors = random_voxels(rs, 2, cfg["sparse_shape"], 700, cfg["in_channels"])
        order = np.lexsort((coors[:, 3], coors[:, 2], coors[:, 1], coors[:, 0]))
        order = order[rs.permutation(len(order))]  # arbitrary voxel order, as voxelization gives
        feats, coors = feats[order], coors[order]
        with torch.no_grad():
            y = enc(torch.from_numpy(feats), torch.from_numpy(coors), 2)
        oy, _ = so.sparse_encoder(specs, feats, coors, 2, cfg["sparse_shape"])
        err = np.abs(oy - y.numpy()).max() / (np.abs(y.numpy()).max() + 1e-12)
        print("encoder %-4s out %s nonzero %d  oracle vs reference rel err %.2e" %
              (tag, tuple(y.shape), int((y != 0).sum()), err))
        assert err < 1e-5
        out["enc_%s_feats" % tag], out["enc_%s_coors" % tag] = feats, coors
        out["enc_%s_out" % tag] = y.numpy()

    # ---- HardSimpleVFE (voxel_encoder.py:29-45): the forward body is two lines of torch ----
    vox = rs.standard_normal((300, 10, 5)).astype(np.float32)
    npts = rs.randint(1, 11, 300).astype(np.int32)
    for i in range(300):
        vox[i, npts[i]:] = 0
    vt, nt = torch.from_numpy(vox), torch.from_numpy(npts)
    mean = vt[:, :, :5].sum(dim=1, keepdim=False) / nt.type_as(vt).view(-1, 1)
    assert np.abs(so.hard_simple_vfe(vox, npts, 5) - mean.numpy()).max() < 1e-6
    out["vfe_voxels"], out["vfe_num"], out["vfe_mean"] = vox, npts, mean.numpy()

    # ---- DynamicVoxelEncoder, plain and virtual ----
    pc_range = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    voxel = [0.8, 0.8, 0.5]
    pts = []
    for b in range(2):
        p = np.zeros((1500, 5), np.float32)
        p[:, :2] = rs.uniform(-53, 53, (1500, 2))
        p[:, 2] = rs.uniform(-5.5, 3.5, 1500)
        p[:, 3:] = rs.uniform(0, 1, (1500, 2))
        p[:40, 0] = 51.2   # points exactly on the (inclusive) upper border
        p[40:60, 1] = -51.2
        pts.append(p)
    enc = dve.DynamicVoxelEncoder(pc_range, voxel, virtual=False)
    v, c, shape = enc([torch.from_numpy(p) for p in pts])
    ov, oc, oshape = so.dynamic_voxel_encoder(pts, pc_range, voxel, False)
    assert np.array_equal(oc, c.numpy()) and np.array_equal(oshape, shape)
    assert np.abs(ov - v.numpy()).max() < 1e-4
    out["dv_range"], out["dv_voxel"] = np.array(pc_range, np.float32), np.array(voxel, np.float32)
    out["dv_pts0"], out["dv_pts1"] = pts
    out["dv_voxels"], out["dv_coors"], out["dv_shape"] = v.numpy(), c.numpy(), shape
    vpts = []
    for b in range(2):
        p = np.zeros((1500, 17), np.float32)
        p[:, :2] = rs.uniform(-53, 53, (1500, 2))
        p[:, 2] = rs.uniform(-5.5, 3.5, 1500)
        p[:, 3:15] = rs.uniform(0, 1, (1500, 12))
        p[:, 15] = rs.choice([1.0, 0.0, -1.0], 1500)
        p[:, 16] = rs.uniform(0, 1, 1500)
        # crowd some voxels so that real and virtual points mix
        p[:400, :3] = p[rs.randint(400, 500, 400), :3] + rs.uniform(-0.1, 0.1, (400, 3)).astype(np.float32)
        vpts.append(p)
    enc = dve.DynamicVoxelEncoder(pc_range, voxel, virtual=True)
    v, c, shape = enc([torch.from_numpy(p) for p in vpts])
    ov, oc, _ = so.dynamic_voxel_encoder(vpts, pc_range, voxel, True)
    assert np.array_equal(oc, c.numpy())
    assert np.abs(ov - v.numpy()).max() < 1e-4, np.abs(ov - v.numpy()).max()
    out["dvv_pts0"], out["dvv_pts1"] = vpts
    out["dvv_voxels"], out["dvv_coors"] = v.numpy(), c.numpy()
    print("dynamic voxel encoder: plain %d voxels, virtual %d voxels (mixed present: %s)" %
          (len(out["dv_voxels"]), len(v), bool(((ov[:, :6] != 0).any(1) & (ov[:, 6:] != 0).any(1)).any())))

    path = os.path.join(GOLDEN, "sparse_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
