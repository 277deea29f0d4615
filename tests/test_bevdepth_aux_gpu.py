"""GPU parity: shift_feature and get_depth_loss (csrc/bevdepth_aux.cu) vs tests/golden/bevdepth_aux.npz,
outputs of the UNMODIFIED reference methods (tools/make_golden_bevdepth.py). Tolerances: warped
features 1e-4 of the range (the sampling coordinates go through a 4x4 inverse), loss 1e-5 rel,
gradients 1e-4 of the largest entry."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "bevdepth_aux.npz"))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_shift_feature(g, cuda):
    x = _t(g["sf_in"], cuda).requires_grad_(True)
    rots = [_t(g["sf_rots0"], cuda), _t(g["sf_rots1"], cuda)]
    trans = [_t(g["sf_trans0"], cuda), _t(g["sf_trans1"], cuda)]
    out = dbev.shift_feature(x, trans, rots, g["sf_dx"], g["sf_bx"])
    want = g["sf_out"]
    assert np.abs(out.detach().cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max()
    (out * _t(g["sf_w"], cuda)).sum().backward()
    wg = g["sf_grad"]
    assert np.abs(x.grad.cpu().numpy() - wg).max() <= 1e-4 * np.abs(wg).max()


def test_shift_feature_identity_and_full_size(cuda):
    """No ego motion -> identity warp; BEVDepth4D size [8, 64, 128, 128] runs and keeps the input."""
    n, v = 8, 6
    rots = torch.eye(3, device=cuda).expand(n, v, 3, 3).contiguous()
    trans = torch.rand(n, v, 3, device=cuda)
    x = torch.randn(n, 64, 128, 128, device=cuda)
    out = dbev.shift_feature(x, [trans, trans], [rots, rots], [0.8, 0.8, 20.0], [-50.8, -50.8, 0.0])
    torch.testing.assert_close(out, x, rtol=1e-4, atol=1e-4)


def test_depth_loss(g, cuda):
    logits = _t(g["dl_logits"], cuda).requires_grad_(True)
    loss = dbev.get_depth_loss(_t(g["dl_gt"], cuda), logits, int(g["dl_D"]), g["dl_dbound"].tolist(),
                               float(g["dl_weight"]))
    np.testing.assert_allclose(float(loss), float(g["dl_loss"]), rtol=1e-5)
    (loss * 2.0).backward()
    wg = 2.0 * g["dl_grad"]
    assert np.abs(logits.grad.cpu().numpy() - wg).max() <= 1e-4 * np.abs(wg).max()


def test_cpu_raises():
    with pytest.raises(RuntimeError):
        dbev.get_depth_loss(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5, 4, 4), 5, [1.0, 6.0, 1.0], 1.0)
