"""GPU parity: fused pillar encoder / scatter / LSS geometry kernels vs reference fixtures
and the numpy oracle. voxel_coors (integers) exact; features within 1e-4 (the reference's
Linear is a cuBLAS/MKL GEMM with unspecified summation order)."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200  # noqa: F401
from distill_bev_b200 import synthetic
from distill_bev_b200.plugin import pillars, view_transformer as vtm
from oracle import lss_oracle, pillar_oracle as po, voxel_oracle as vo

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _load_encoder(g, dev, vs, pcr):
    enc = pillars.DynamicPillarFeatureNet(in_channels=g["points"].shape[1], feat_channels=(64,),
                                          voxel_size=vs, point_cloud_range=pcr,
                                          norm_cfg=dict(type="BN1d", eps=float(g["bn_eps"]), momentum=0.01)).to(dev)
    with torch.no_grad():
        enc.pfn_layers[0][0].weight.copy_(_t(g["weight"], dev))
        bn = enc.pfn_layers[0][1]
        bn.weight.copy_(_t(g["bn_weight"], dev)); bn.bias.copy_(_t(g["bn_bias"], dev))
        bn.running_mean.copy_(_t(g["bn_mean"], dev)); bn.running_var.copy_(_t(g["bn_var"], dev))
    return enc


def test_pillar_encoder_golden(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "pillar_small.npz"))
    vs, pcr = g["voxel_size"].tolist(), g["coors_range"].tolist()
    enc = _load_encoder(g, cuda, vs, pcr).eval()
    # state_dict keys are the reference's (checkpoints load unchanged)
    assert set(enc.state_dict()) >= {"pfn_layers.0.0.weight", "pfn_layers.0.1.weight", "pfn_layers.0.1.bias",
                                     "pfn_layers.0.1.running_mean", "pfn_layers.0.1.running_var"}
    vf, vc = enc(_t(g["points"], cuda), _t(g["coors"], cuda))
    np.testing.assert_array_equal(vc.cpu().numpy(), g["voxel_coors"])
    np.testing.assert_allclose(vf.cpu().numpy(), g["voxel_feats"], rtol=1e-4, atol=1e-4)
    for cl in (False, True):
        canvas = pillars.PointPillarsScatter(64, [64, 64], channels_last=cl)(vf, vc, 2)
        assert tuple(canvas.shape) == (2, 64, 64, 64)
        c = canvas.cpu().numpy()
        np.testing.assert_array_equal(np.stack(np.nonzero(c.sum(1))).astype(np.int32), g["canvas_nonzero"])
        np.testing.assert_array_equal(c, po.pillar_scatter(vf.cpu().numpy(), vc.cpu().numpy(), 2, 64, 64))
    # training-mode composition (batch-norm batch statistics) == torch reference of the same flow
    enc.train()
    vf2, vc2 = enc(_t(g["points"], cuda), _t(g["coors"], cuda))
    np.testing.assert_array_equal(vc2.cpu().numpy(), g["voxel_coors"])
    assert torch.isfinite(vf2).all() and vf2.shape == vf.shape


@pytest.mark.parametrize("B,n", [(8, 30000), (2, 240000)])
def test_pillar_canvas_full_size_vs_oracle(cuda, B, n):
    """configs[1] teacher input: B clouds of n points, 0.2 m pillars, 512x512 canvas."""
    vs, pcr = [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    clouds = synthetic.make_lidar(B, n, seed=2)
    torch.manual_seed(0)
    enc = pillars.DynamicPillarFeatureNet(in_channels=5, feat_channels=(64,), voxel_size=vs,
                                          point_cloud_range=pcr).to(cuda).eval()
    bn = enc.pfn_layers[0][1]
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5)
    scat = pillars.PointPillarsScatter(64, [512, 512], channels_last=True)
    canvas = pillars.pillar_canvas([_t(c, cuda) for c in clouds], enc, scat)
    assert tuple(canvas.shape) == (B, 64, 512, 512)
    coors = np.concatenate([np.concatenate([np.full((n, 1), b, np.int32), vo.dynamic_voxelize(c, vs, pcr)], 1)
                            for b, c in enumerate(clouds)])
    vf, vc = po.pillar_encode(np.concatenate(clouds), coors, enc.pfn_layers[0][0].weight.detach().cpu().numpy(),
                              bn.weight.detach().cpu().numpy(), bn.bias.detach().cpu().numpy(),
                              bn.running_mean.cpu().numpy(), bn.running_var.cpu().numpy(), bn.eps, vs, pcr)
    ref = po.pillar_scatter(vf, vc, B, 512, 512)
    got = canvas.cpu().numpy()
    np.testing.assert_array_equal(got.sum(1) != 0, ref.sum(1) != 0)      # same occupied pillars
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("nout,nfeat", [(64, 5), (32, 4), (20, 4)])
def test_pillar_canvas_equals_encode_then_scatter(cuda, channels_last, nout, nfeat):
    """The fused canvas entry point stores the same bits as dbev_pillar_encode + dbev_pillar_scatter."""
    vs, pcr = [0.4, 0.4, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    clouds = [c[:, :nfeat].copy() for c in synthetic.make_lidar(3, 20000, seed=4)]
    torch.manual_seed(nout)
    enc = pillars.DynamicPillarFeatureNet(in_channels=nfeat, feat_channels=(nout,), voxel_size=vs,
                                          point_cloud_range=pcr).to(cuda).eval()
    scat = pillars.PointPillarsScatter(nout, [256, 256], channels_last=channels_last)
    pts = [_t(c, cuda) for c in clouds]
    canvas = pillars.pillar_canvas(pts, enc, scat)
    lin, bn = enc.pfn_layers[0][0], enc.pfn_layers[0][1]
    scale, shift = pillars.fold_bn(bn)
    offs = torch.tensor([0, 20000, 40000, 60000], dtype=torch.int32)
    vf, vc = pillars.pillar_encode(torch.cat(pts), lin.weight, scale, shift, vs, pcr, 3, batch_offsets=offs)
    two_step = pillars.pillar_scatter(vf, vc, 3, 256, 256, channels_last)
    assert canvas.shape == two_step.shape and canvas.stride() == two_step.stride()
    assert torch.equal(canvas, two_step)
    assert int((canvas != 0).any(1).sum()) <= vf.shape[0]


def test_lss_geometry_vs_reference_golden(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "lss_small.npz"))
    geom = vtm.lss_geometry(_t(g["frustum"], cuda), _t(g["rots"], cuda), _t(g["trans"], cuda),
                            _t(g["intrins"], cuda), _t(g["post_rots"], cuda), _t(g["post_trans"], cuda))
    np.testing.assert_allclose(geom.cpu().numpy(), g["geom"], rtol=2e-5, atol=2e-4)


def test_view_transformer_module_config1(cuda):
    """ViewTransformerLiftSplatShoot mirror: attributes, geometry and pooling on configs[0]."""
    vt = vtm.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=32, numC_Trans=64).to(cuda)
    assert vt.D == 59 and tuple(vt.frustum.shape) == (59, 16, 44, 3)
    # torch.linspace on the CPU is 1 ulp platform dependent (vectorised FMA vs scalar path)
    np.testing.assert_allclose(vt.frustum.cpu().numpy(),
                               lss_oracle.create_frustum((256, 704), 16, synthetic.NUSC_GRID["dbound"]),
                               rtol=3e-7, atol=0)
    calib = synthetic.make_calibration(1, 6, seed=0)
    geom = vt.get_geometry(*[_t(a, cuda) for a in calib])
    ref_geom = lss_oracle.get_geometry(vt.frustum.cpu().numpy(), *calib)
    np.testing.assert_allclose(geom.cpu().numpy(), ref_geom, rtol=2e-5, atol=5e-4)
    # cell assignment of the device geometry vs the oracle geometry: only border points may move
    i1, k1 = lss_oracle.voxel_indices(geom.cpu().numpy(), vt.bx.cpu().numpy(), vt.dx.cpu().numpy(), vt.nx.cpu().numpy())
    i2, k2 = lss_oracle.voxel_indices(ref_geom, vt.bx.cpu().numpy(), vt.dx.cpu().numpy(), vt.nx.cpu().numpy())
    moved = int((k1 != k2).sum() + ((i1 != i2).any(1) & k1 & k2).sum())
    assert moved <= 0.0005 * k2.size, moved
    x = torch.rand(1, 6, 59, 16, 44, 64, device=cuda)
    out = vt.voxel_pooling(geom, x)
    ref = lss_oracle.voxel_pooling(geom.cpu().numpy(), x.cpu().numpy(), vt.bx.cpu().numpy(), vt.dx.cpu().numpy(),
                                   vt.nx.cpu().numpy())
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    feats = torch.rand(1, 6, 32, 16, 44, device=cuda)
    bev = vt((feats,) + tuple(_t(a, cuda) for a in calib))
    assert tuple(bev.shape) == (1, 64, 128, 128) and torch.isfinite(bev).all()


def test_bevdepth_view_transformer_forward_backward(cuda):
    """ViewTransformerLSSBEVDepth.forward == the same sub-modules composed like the reference's forward
    (view_transformer_mine.py:312-344: lift volume + voxel_pooling), without building the volume; gradients reach the
    depth branch (se / extra_depthnet / dcn / depthnet) and featnet."""
    import distill_bev_b200 as dbev
    from distill_bev_b200 import synthetic
    torch.manual_seed(0)
    vt = dbev.ViewTransformerLSSBEVDepth(
        extra_depth_net=dict(type='ResNetForBEVDet', numC_input=256, num_layer=[1, ], num_channels=[256, ], stride=[1, ]),
        loss_depth_weight=100.0, grid_config=synthetic.NUSC_GRID, numC_input=128, numC_Trans=64).to(cuda).train()
    B, N = 1, 6
    calib = [torch.from_numpy(a).to(cuda) for a in synthetic.make_calibration(B, N, seed=2)]
    rots, trans, intrins, post_rots, post_trans = [c.view(B, N, *c.shape[2:]) for c in calib]
    x = torch.randn(B, N, 128, 16, 44, device=cuda, requires_grad=True)
    bev, depth_digit = vt((x, rots, trans, intrins, post_rots, post_trans, None))
    assert tuple(bev.shape) == (B, 64, 128, 128) and tuple(depth_digit.shape) == (B * N, 59, 16, 44)
    with torch.no_grad():
        xf = x.view(B * N, 128, 16, 44)
        img_feat = vt.featnet(xf)
        depth = vt.get_depth_dist(depth_digit)
        volume = (depth.unsqueeze(1) * img_feat.unsqueeze(2)).view(B, N, 64, 59, 16, 44).permute(0, 1, 3, 4, 5, 2)
        want = vt.voxel_pooling(vt.get_geometry(rots, trans, intrins, post_rots, post_trans), volume.contiguous())
    torch.testing.assert_close(bev.detach(), want, rtol=1e-4, atol=1e-5 * float(want.abs().max()))
    (bev.sum() + depth_digit.sum()).backward()
    for p in (vt.featnet.weight, vt.depthnet.weight, vt.se.input_conv.weight, vt.dcn[0].weight,
              vt.extra_depthnet.layers[0][0].conv1.weight):
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0
    assert torch.isfinite(x.grad).all()
