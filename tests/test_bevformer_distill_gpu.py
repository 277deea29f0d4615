"""GPU parity: BEVFormer-student distillation variants (plugin/distill/bevformer.py) vs
tests/golden/bevformer_small.npz — outputs of the UNMODIFIED BEVFormerDistill methods
(tools/make_golden_bevformer.py). Masks / counts exact; losses rtol 1e-4; gradients 1e-4 of max."""
import json
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev
from distill_bev_b200.plugin.distill import bevformer as bf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "bevformer_small.npz"))


def _split(flat, counts):
    out, o = [], 0
    for c in counts:
        out.append(torch.from_numpy(flat[o:o + c].copy()))
        o += c
    return out


def _cfg(g):
    return dict(grid_size=g["grid"].tolist(), point_cloud_range=g["pc_range"].tolist(),
                voxel_size=g["voxel"].tolist())


def test_masks(g, cuda):
    H = g["teacher"].shape[-1]
    gt = _split(g["gt"], g["n_gt"])
    fg, fgs, bgs = bf.foreground_scale_mask(H, H, gt, _cfg(g), cuda)
    assert np.array_equal(fg.cpu().numpy(), g["recipe_fg"])
    np.testing.assert_allclose(fgs.cpu().numpy(), g["recipe_fg_scale"], rtol=1e-5)
    np.testing.assert_allclose(bgs.cpu().numpy(), g["recipe_bg_scale"], rtol=1e-6)
    preds = [(p, s, None) for p, s in zip(_split(g["pred"], g["n_pred"]), _split(g["scores"], g["n_pred"]))]
    params = json.loads(str(g["recipe_params"]))
    fp, fps, cnt = bf.add_fp_as_fg_bbox(H, H, "teacher", fg, preds, gt, params, _cfg(g))
    assert np.array_equal(fp.cpu().numpy(), g["recipe_fp"])
    assert np.array_equal(cnt.cpu().numpy(), g["recipe_fp_count"])
    np.testing.assert_allclose(fps.cpu().numpy(), g["recipe_fp_scale"], rtol=1e-6)


@pytest.mark.parametrize("name", ["recipe", "nofp"])
def test_fgd_loss_and_grad(g, cuda, name):
    params = json.loads(str(g[name + "_params"]))
    H = g["teacher"].shape[-1]
    gt = _split(g["gt"], g["n_gt"])
    preds = [(p, s, None) for p, s in zip(_split(g["pred"], g["n_pred"]), _split(g["scores"], g["n_pred"]))]
    conv = torch.nn.Conv2d(1, 1, 3, padding=1).to(cuda)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(g[name + "_conv_w"]).view(1, 1, 3, 3))
        conv.bias.copy_(torch.from_numpy(g[name + "_conv_b"]))
    st = torch.from_numpy(g["student"]).to(cuda).requires_grad_(True)
    out = bf.fgd_distill_loss(torch.from_numpy(g["teacher"]).to(cuda), st, gt, preds, params, _cfg(g),
                              spatial_adaptation=conv, epoch=5)
    keys = json.loads(str(g[name + "_loss_keys"]))
    assert sorted(out) == keys
    got = np.array([float(out[k]) for k in keys])
    np.testing.assert_allclose(got, g[name + "_loss_vals"], rtol=1e-4)
    sum(out.values()).backward()
    want = g[name + "_grad_student"]
    assert np.abs(st.grad.cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max()


def test_hs_and_query_losses(g, cuda):
    params = json.loads(str(g["recipe_params"]))
    t_hs, s_hs = torch.from_numpy(g["t_hs"]).to(cuda), torch.from_numpy(g["s_hs"]).to(cuda)
    hs = bf.hs_distill_loss(t_hs[-1].permute(0, 2, 1), s_hs[-1].permute(0, 2, 1), params)
    np.testing.assert_allclose(float(hs["hs_feat_loss"]), float(g["hs_loss"]), rtol=1e-5)
    q = bf.query_distill_loss(torch.from_numpy(g["teacher"]).to(cuda), torch.from_numpy(g["t_query"]).to(cuda),
                              t_hs, torch.from_numpy(g["student"]).to(cuda),
                              torch.from_numpy(g["s_query"]).to(cuda), s_hs, params)
    np.testing.assert_allclose(float(q["query_loss"]), float(g["query_loss"]), rtol=1e-4)
