"""GPU parity: affinity distillation loss (csrc/affinity.cu) vs the oracle restatement of
BEVDetDistill.affinity_distill_loss (pinned by tests/golden/fgd_small.npz, produced by the
unmodified reference method) and vs a float64 PyTorch autograd reference for the gradient.
Tolerances: loss rtol 1e-4, gradient 1e-4 of the largest entry."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev
from oracle import fgd_oracle as fo

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _rows(feat, mask):
    """feat [C,H,W], mask [H,W] -> [K, C] in row-major cell order (bevdet_distill.py:1317-1320)."""
    return feat[:, mask].T


def _torch_ref(t, s, mask, weight, kind, beta=1.0):
    t64 = torch.from_numpy(t).double()
    s64 = torch.from_numpy(s).double().requires_grad_(True)
    total = 0
    for b in range(t.shape[0]):
        m = torch.from_numpy(mask[b].astype(bool))
        tr, sr = t64[b][:, m].T, s64[b][:, m].T
        if tr.shape[0] == 0:
            continue
        ta, sa = tr @ tr.T, sr @ sr.T
        if kind == "SmoothL1Loss":
            l = torch.nn.functional.smooth_l1_loss(ta, sa, beta=beta)
        elif kind == "L1Loss":
            l = torch.nn.functional.l1_loss(ta, sa)
        else:
            l = torch.nn.functional.mse_loss(ta, sa)
        total = total + l * weight
    total.backward()
    return float(total), s64.grad.numpy()


def test_golden_rows(golden_dir, cuda):
    g = np.load(os.path.join(golden_dir, "fgd_small.npz"))
    t_rows = [g["aff_t0"], g["aff_t1"]]
    s_rows = [g["aff_s0"], g["aff_s1"]]
    C = t_rows[0].shape[1]
    # place the rows into a [B, C, 1, K] map with an all-ones mask of the right length
    K = max(r.shape[0] for r in t_rows)
    t = np.zeros((2, C, 1, K), np.float32)
    s = np.zeros((2, C, 1, K), np.float32)
    mask = np.zeros((2, 1, 1, K), np.float32)
    for b in range(2):
        k = t_rows[b].shape[0]
        t[b, :, 0, :k], s[b, :, 0, :k], mask[b, 0, 0, :k] = t_rows[b].T, s_rows[b].T, 1
    out = dbev.affinity.affinity_distill_loss(_t(t, cuda), _t(s, cuda), _t(mask, cuda), weight=0.5)
    np.testing.assert_allclose(float(out["kd_affinity_loss"]), float(g["aff_loss"]), rtol=1e-4)


@pytest.mark.parametrize("kind", ["SmoothL1Loss", "L1Loss", "MSELoss"])
def test_loss_and_grad(cuda, kind):
    rs = np.random.RandomState(1)
    B, C, H, W = 3, 96, 24, 20
    t = (rs.standard_normal((B, C, H, W)) * 0.2).astype(np.float32)
    s = (rs.standard_normal((B, C, H, W)) * 0.2).astype(np.float32)
    mask = (rs.uniform(size=(B, H, W)) < 0.35)
    mask[1] = False                      # a sample without selected cells
    mask[2, :3] = True
    want, want_grad = _torch_ref(t, s, mask, 0.7, kind)
    oracle = fo.affinity_loss([_rows(t[b], mask[b]) for b in range(B) if mask[b].any()],
                              [_rows(s[b], mask[b]) for b in range(B) if mask[b].any()], 0.7) \
        if kind == "SmoothL1Loss" else want
    st = _t(s, cuda).requires_grad_(True)
    out = dbev.affinity.affinity_distill_loss(_t(t, cuda), st, _t(mask[:, None].astype(np.float32), cuda),
                                              weight=0.7, criterion=dict(type=kind))
    loss = out["kd_affinity_loss"]
    np.testing.assert_allclose(float(loss), want, rtol=1e-4)
    np.testing.assert_allclose(float(loss), oracle, rtol=1e-4)
    (loss * 1.5).backward()
    got = st.grad.cpu().numpy()
    assert np.abs(got - 1.5 * want_grad).max() <= 1e-4 * np.abs(want_grad).max() * 1.5 + 1e-9
    assert (got[1] == 0).all()


def test_two_masks_and_channel_tail(cuda):
    """'foreground+fp' (two masks OR-ed, :1296-1299) and a channel count that is not a multiple of
    the 128-channel backward chunk."""
    rs = np.random.RandomState(2)
    B, C, H, W = 2, 200, 16, 16
    t = (rs.standard_normal((B, C, H, W)) * 0.15).astype(np.float32)
    s = (rs.standard_normal((B, C, H, W)) * 0.15).astype(np.float32)
    ma = rs.uniform(size=(B, H, W)) < 0.2
    mb = rs.uniform(size=(B, H, W)) < 0.2
    want, want_grad = _torch_ref(t, s, ma | mb, 1.0, "SmoothL1Loss")
    st = _t(s, cuda).requires_grad_(True)
    out = dbev.affinity.affinity_distill_loss(_t(t, cuda), st, _t(ma[:, None].astype(np.float32), cuda),
                                              _t(mb[:, None].astype(np.float32), cuda))
    np.testing.assert_allclose(float(out["kd_affinity_loss"]), want, rtol=1e-4)
    out["kd_affinity_loss"].backward()
    assert np.abs(st.grad.cpu().numpy() - want_grad).max() <= 1e-4 * np.abs(want_grad).max()


def test_large_k_deterministic(cuda):
    """K ~ 1000 rows per sample at 384 channels (affinity_attention_topk=1000 scale): run twice,
    bit-identical (no float atomics), and loss equal to the fp64 reference."""
    rs = np.random.RandomState(4)
    B, C, H, W = 2, 384, 64, 64
    t = (rs.standard_normal((B, C, H, W)) * 0.05).astype(np.float32)
    s = (rs.standard_normal((B, C, H, W)) * 0.05).astype(np.float32)
    mask = rs.uniform(size=(B, H, W)) < 0.25
    want, _ = _torch_ref(t, s, mask, 1.0, "SmoothL1Loss")
    tt, ss, mm = _t(t, cuda), _t(s, cuda), _t(mask[:, None].astype(np.float32), cuda)
    a = dbev.affinity.affinity_distill_loss(tt, ss, mm)["kd_affinity_loss"]
    b = dbev.affinity.affinity_distill_loss(tt, ss, mm)["kd_affinity_loss"]
    assert torch.equal(a, b)
    np.testing.assert_allclose(float(a), want, rtol=1e-4)


@pytest.mark.parametrize("split", [2, 3])
def test_affinity_split_partitions(cuda, split):
    """affinity_split > 1 (:738-747): rows of a sample partitioned by a permutation into perm[j::split], one gram pair
    per part, averaged. Explicit permutations vs the oracle; default permutation = torch.randperm per sample from the
    CPU generator (the reference's own draw), checked by re-seeding."""
    rng = np.random.RandomState(split)
    B, C, H = 3, 16, 12
    t = rng.randn(B, C, H, H).astype(np.float32)
    s = rng.randn(B, C, H, H).astype(np.float32)
    mask = (rng.rand(B, 1, H, H) < 0.2).astype(np.float32)
    mask[2] = 0
    mask[2, 0, 0, :2] = 1                                     # 2 rows: with split 3 one part is empty
    ks = [int(mask[b].sum()) for b in range(B)]
    perms = [rng.permutation(k) for k in ks]
    t_rows = [_rows(t[b], mask[b, 0] > 0) for b in range(B)]
    s_rows = [_rows(s[b], mask[b, 0] > 0) for b in range(B)]
    want = fo.affinity_loss(t_rows, s_rows, 0.7, perm=perms, split=split)
    st = _t(s, cuda).requires_grad_(True)
    got = dbev.affinity.affinity_distill_loss(_t(t, cuda), st, _t(mask, cuda), weight=0.7, split=split, perms=perms)["kd_affinity_loss"]
    assert abs(float(got) - want) <= 1e-4 * abs(want), (float(got), want)
    got.backward()
    assert torch.isfinite(st.grad).all() and float(st.grad.abs().sum()) > 0
    # default permutation: the same torch.randperm draws as the reference, in sample order
    torch.manual_seed(123)
    a = dbev.affinity.affinity_distill_loss(_t(t, cuda), _t(s, cuda), _t(mask, cuda), weight=0.7, split=split)["kd_affinity_loss"]
    torch.manual_seed(123)
    ref_perms = [torch.randperm(k).numpy() for k in ks]
    want2 = fo.affinity_loss(t_rows, s_rows, 0.7, perm=ref_perms, split=split)
    assert abs(float(a) - want2) <= 1e-4 * abs(want2), (float(a), want2)
