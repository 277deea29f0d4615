"""GPU microbench of the sparse teacher path at LidarFormer size (41 x 1600 x 1600 grid,
configs/teacher_transformer/lidarformer.py:43-51): rulebooks, every conv of the encoder, dense.
Prints one JSON object; per-layer times come from CUDA events around each call."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200 import synthetic  # noqa: E402
from distill_bev_b200.plugin.ops import spconv as sp  # noqa: E402

LF = dict(in_channels=5, sparse_shape=[41, 1600, 1600], output_channels=128,
          encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
          encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)),
          block_type="basicblock")


def make_voxels(batch, n_points, dev, seed=0):
    """Hard-voxelize synthetic clouds with the LidarFormer voxel size -> mean VFE features."""
    vox = dbev.Voxelization([0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 10, (90000, 120000)).eval()
    vfe = dbev.HardSimpleVFE(5)
    feats, coors = [], []
    for b, pts in enumerate(synthetic.make_lidar_scene(batch, n_points, seed=seed)):
        v, c, n = vox(torch.from_numpy(pts).to(dev))
        feats.append(vfe(v, n, c))
        coors.append(torch.nn.functional.pad(c, (1, 0), value=b))
    return torch.cat(feats), torch.cat(coors).contiguous()


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2], r


def reference_kernels(enc, feats, coors, batch, dev):
    """The kernels to beat: the reference's own spconv CUDA extension (mmdet3d/ops/spconv, compiled
    unmodified for sm_100 into oracle/_ref by oracle/build_oracle.py) on the same inputs — rulebook
    (get_indice_pairs_3d) and indice_conv_fp32 of the first SubM layers, a strided layer and a
    128-channel SubM layer. Skipped when the extension is not in the snapshot."""
    import glob
    import importlib.util
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_sparse_conv_ext*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("ref_sparse_conv_ext", so[0])
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    out = []
    shape = LF["sparse_shape"]

    def pairs(indices, in_shape, k, s, p, subm):
        o_shape = in_shape if subm else sp.get_conv_output_size(in_shape, k, s, p, [1, 1, 1])
        fn = lambda: ext.get_indice_pairs_3d(indices, batch, o_shape, in_shape, k, s, p, [1, 1, 1],
                                             [0, 0, 0], int(subm), 0)
        ms, r = timed(fn, 3)
        return ms, r, o_shape

    def conv(x, cin, cout, r, subm, n_out):
        w = torch.randn(3, 3, 3, cin, cout, device=dev) * 0.05
        fn = lambda: ext.indice_conv_fp32(x, w, r[1], r[2], n_out, 0, int(subm))
        ms, y = timed(fn, 3)
        return ms, y

    ms_rb, r, _ = pairs(coors, shape, [3, 3, 3], [1, 1, 1], [1, 1, 1], True)
    x16 = torch.randn(coors.shape[0], 16, device=dev)
    ms_c, _ = conv(x16, 16, 16, r, True, coors.shape[0])
    out.append(dict(layer="subm 16->16 stage 1", n_out=int(coors.shape[0]), rulebook_ms=ms_rb, conv_ms=ms_c))
    ms_rb2, r2, s2 = pairs(coors, shape, [3, 3, 3], [2, 2, 2], [1, 1, 1], False)
    n2 = int(r2[0].shape[0])
    ms_c2, _ = conv(x16, 16, 32, r2, False, n2)
    out.append(dict(layer="sparse 16->32 stride 2", n_out=n2, rulebook_ms=ms_rb2, conv_ms=ms_c2))
    # deeper stages: downsample the coordinates the way the encoder does to get realistic sets
    ind = r2[0]
    for cin, cout in ((32, 64), (64, 128)):
        ms_rb3, r3, s3 = pairs(ind, s2, [3, 3, 3], [2, 2, 2], [1, 1, 1], False)
        ind, s2 = r3[0], s3
    ms_rb4, r4, _ = pairs(ind, s2, [3, 3, 3], [1, 1, 1], [1, 1, 1], True)
    x128 = torch.randn(ind.shape[0], 128, device=dev)
    ms_c4, _ = conv(x128, 128, 128, r4, True, ind.shape[0])
    out.append(dict(layer="subm 128->128 stage 4", n_out=int(ind.shape[0]), rulebook_ms=ms_rb4, conv_ms=ms_c4))
    return out


def main(batch=4, n_points=240000, impl="auto"):
    dev = torch.device("cuda:0")
    feats, coors = make_voxels(batch, n_points, dev)
    enc = dbev.SparseEncoder(**LF).to(dev).eval()
    sp.SparseConvolution.impl = None if impl == "auto" else impl
    res = dict(batch=batch, n_points=n_points, n_voxels=int(feats.shape[0]), impl=impl)
    t, out = timed(lambda: enc(feats, coors, batch))
    res["encoder_ms"] = t
    res["out_shape"] = list(out.shape)
    # per-layer breakdown: hook every conv's rulebook + kernel
    layers = []
    orig_build, orig_conv = sp.build_rulebook, sp.conv_table

    def build(*a, **k):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); rb = orig_build(*a, **k); e1.record(); torch.cuda.synchronize()
        layers.append(dict(op="rulebook", subm=bool(a[-1]), n_in=int(a[0].shape[0]), n_out=int(rb.n_out),
                           ms=e0.elapsed_time(e1)))
        return rb

    def conv(features, weight, nbr, n_out, *a, **k):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); y = orig_conv(features, weight, nbr, n_out, *a, **k); e1.record(); torch.cuda.synchronize()
        pairs = int((nbr >= 0).sum())
        cin, cout = weight.shape[-2], weight.shape[-1]
        ms = e0.elapsed_time(e1)
        layers.append(dict(op="conv", cin=cin, cout=cout, kvol=int(nbr.shape[0]), n_out=int(n_out), pairs=pairs,
                           ms=ms, gflops=2.0 * pairs * cin * cout / ms / 1e6,
                           gbs=(pairs * cin + n_out * cout) * 4 / ms / 1e6))
        return y

    sp.build_rulebook, sp.conv_table = build, conv
    try:
        enc(feats, coors, batch)
    finally:
        sp.build_rulebook, sp.conv_table = orig_build, orig_conv
    res["layers"] = layers
    ref = reference_kernels(enc, feats, coors, batch, dev)
    if ref:
        res["reference_cuda_ext"] = ref
    res["sum_rulebook_ms"] = sum(l["ms"] for l in layers if l["op"] == "rulebook")
    res["sum_conv_ms"] = sum(l["ms"] for l in layers if l["op"] == "conv")
    print(json.dumps(res))


if __name__ == "__main__":
    a = sys.argv[1:]
    main(int(a[0]) if a else 4, int(a[1]) if len(a) > 1 else 240000, a[2] if len(a) > 2 else "auto")
