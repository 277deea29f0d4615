"""Hard-voxel PillarFeatureNet (pillar_encoder.py:14-162): the numpy oracle and the fused sm_100a
kernel against the fixture generated from the UNMODIFIED reference class (tools/make_golden_pillar_hard.py), eval mode,
legacy False (the shipped config) and True."""
import os

import numpy as np
import pytest
import torch

from oracle import pillar_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pillar_hard.npz")


def _args(gd):
    return dict(weight=gd["sd/pfn_layers.0.linear.weight"], bn_weight=gd["sd/pfn_layers.0.norm.weight"],
                bn_bias=gd["sd/pfn_layers.0.norm.bias"], bn_mean=gd["sd/pfn_layers.0.norm.running_mean"],
                bn_var=gd["sd/pfn_layers.0.norm.running_var"], bn_eps=1e-3, voxel_size=list(gd["voxel_size"]),
                point_cloud_range=list(gd["range"]))


@pytest.mark.parametrize("legacy", [False, True])
def test_oracle_matches_reference_class(legacy):
    gd = np.load(GOLDEN)
    got = pillar_oracle.hard_pillar_encode(gd["voxels"], gd["num_points"], gd["coors"], legacy=legacy, **_args(gd))
    want = gd["out_legacy" if legacy else "out_new"]
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_state_dict_keys_match_reference_class():
    import distill_bev_b200 as dbev
    gd = np.load(GOLDEN)
    net = dbev.PillarFeatureNet(in_channels=5, feat_channels=[64], voxel_size=(0.2, 0.2, 8),
                                point_cloud_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), legacy=False)
    assert list(net.state_dict().keys()) == [str(k) for k in gd["keys"]]
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == tuple(gd["sd/" + k].shape), k
    with pytest.raises(NotImplementedError):
        dbev.PillarFeatureNet(in_channels=5, feat_channels=[64, 64])


@pytest.mark.gpu
@pytest.mark.parametrize("legacy", [False, True])
def test_kernel_matches_reference_class(cuda, legacy):
    import distill_bev_b200 as dbev
    gd = np.load(GOLDEN)
    net = dbev.PillarFeatureNet(in_channels=5, feat_channels=[64], voxel_size=tuple(gd["voxel_size"]),
                                point_cloud_range=tuple(gd["range"]), legacy=legacy)
    net.load_state_dict({str(k): torch.from_numpy(gd["sd/" + str(k)]) for k in gd["keys"]}, strict=True)
    net = net.to(cuda).eval()
    out = net(torch.from_numpy(gd["voxels"]).to(cuda), torch.from_numpy(gd["num_points"]).to(cuda),
              torch.from_numpy(gd["coors"]).to(cuda))
    want = torch.from_numpy(gd["out_legacy" if legacy else "out_new"]).to(cuda)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    with pytest.raises(NotImplementedError):
        net.train()(torch.zeros(1, 20, 5, device=cuda), torch.ones(1, device=cuda), torch.zeros(1, 4, device=cuda))


@pytest.mark.gpu
def test_kernel_full_size_teacher_front_end(cuda):
    """8 x 30k-point clouds through Voxelization (hard, max 20 points, the teacher's eval max_voxels) ->
    PillarFeatureNet -> PointPillarsScatter: the kernel equals the oracle on every pillar; the canvas equals the
    scattered oracle rows."""
    import distill_bev_b200 as dbev
    from distill_bev_b200 import synthetic
    vs, rng = [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    vox = dbev.Voxelization(vs, rng, 20, (30000, 40000)).eval()
    torch.manual_seed(0)
    net = dbev.PillarFeatureNet(in_channels=5, feat_channels=[64], voxel_size=tuple(vs), point_cloud_range=tuple(rng),
                                legacy=False).to(cuda).eval()
    bn = net.pfn_layers[0].norm
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.3), bn.running_var.uniform_(0.5, 1.5), bn.weight.uniform_(0.5, 1.5), bn.bias.normal_(0, 0.3)
    vs_l, ns_l, cs_l = [], [], []
    for b, pts in enumerate(synthetic.make_lidar(2, 30000, seed=4)):
        v, c, n = vox(torch.from_numpy(pts).to(cuda))
        vs_l.append(v), ns_l.append(n), cs_l.append(torch.nn.functional.pad(c, (1, 0), value=b))
    voxels, num_points, coors = torch.cat(vs_l), torch.cat(ns_l), torch.cat(cs_l)
    out = net(voxels, num_points, coors)
    want = pillar_oracle.hard_pillar_encode(
        voxels.cpu().numpy(), num_points.cpu().numpy(), coors.cpu().numpy(), net.pfn_layers[0].linear.weight.detach().cpu().numpy(),
        bn.weight.detach().cpu().numpy(), bn.bias.detach().cpu().numpy(), bn.running_mean.cpu().numpy(),
        bn.running_var.cpu().numpy(), bn.eps, vs, rng, legacy=False)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    canvas = dbev.PointPillarsScatter(64, [512, 512])(out, coors, 2)
    ref = pillar_oracle.pillar_scatter(want, coors.cpu().numpy(), 2, 512, 512)
    np.testing.assert_allclose(canvas.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_virtual_points_match_reference_class(cuda):
    """virtual=True (MVP teachers, pillar_encoder.py:108-113): the virtual label column (-1 = virtual) is rewritten to
    1 / 0 in place before the decorations, like the reference; output against the UNMODIFIED class on the same voxels."""
    import distill_bev_b200 as dbev
    gd = np.load(GOLDEN)
    vmask = np.unpackbits(gd["virtual_mask"], axis=1)[:, :gd["voxels"].shape[1]].astype(bool)
    feats = torch.from_numpy(gd["voxels"]).clone()
    feats[..., -2][torch.from_numpy(vmask)] = -1.0
    net = dbev.PillarFeatureNet(in_channels=5, feat_channels=[64], voxel_size=tuple(gd["voxel_size"]),
                                point_cloud_range=tuple(gd["range"]), legacy=False, virtual=True)
    net.load_state_dict({str(k): torch.from_numpy(gd["sd/" + str(k)]) for k in gd["keys"]}, strict=True)
    net = net.to(cuda).eval()
    f = feats.to(cuda)
    out = net(f, torch.from_numpy(gd["num_points"]).to(cuda), torch.from_numpy(gd["coors"]).to(cuda))
    torch.testing.assert_close(out, torch.from_numpy(gd["out_virtual"]).to(cuda), rtol=1e-5, atol=1e-5)
    assert torch.equal(f[..., -2].cpu(), torch.from_numpy(vmask).float())          # rewritten in place, as the reference does
    # the dynamic-voxel encoder applies the same rewrite (pillar_encoder.py:294-299)
    dyn = dbev.DynamicPillarFeatureNet(in_channels=5, feat_channels=(64,), voxel_size=(0.2, 0.2, 8),
                                       point_cloud_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), virtual=True).to(cuda).eval()
    plain = dbev.DynamicPillarFeatureNet(in_channels=5, feat_channels=(64,), voxel_size=(0.2, 0.2, 8),
                                         point_cloud_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)).to(cuda).eval()
    plain.load_state_dict(dyn.state_dict())
    from distill_bev_b200 import synthetic
    pts = torch.from_numpy(synthetic.make_lidar(1, 4000, seed=5)[0]).to(cuda)
    lab = torch.rand(pts.shape[0], device=cuda) < 0.3
    raw = pts.clone()
    raw[:, -2] = torch.where(lab, torch.full_like(raw[:, -2], -1.0), raw[:, -2].abs() + 0.5)
    pre = raw.clone()
    pre[:, -2] = lab.float()
    vox = dbev.Voxelization([0.2, 0.2, 8], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], -1, -1)
    co = torch.nn.functional.pad(vox(pts), (1, 0), value=0)
    a, ca = dyn(raw, co)
    b, cb = plain(pre, co)
    assert torch.equal(a, b) and torch.equal(ca, cb)
