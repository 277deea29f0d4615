"""GPU parity: multi-scale deformable attention (csrc/ms_deform_attn.cu) vs a plain PyTorch fp32/fp64
restatement of mmcv's published multi_scale_deformable_attn_pytorch (F.grid_sample bilinear, zeros padding,
align_corners=False per level; third party -> parity unpinned by reference fixtures). Forward 1e-5,
gradients 1e-4 of the largest entry (float atomics in grad_value)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import distill_bev_b200 as dbev

pytestmark = pytest.mark.gpu


def reference(value, shapes, loc, attn):
    bs, _, heads, dim = value.shape
    _, nq, _, L, P, _ = loc.shape
    vals = value.split([int(h * w) for h, w in shapes], dim=1)
    grids = 2 * loc - 1
    out = []
    for l, (h, w) in enumerate(shapes):
        v = vals[l].flatten(2).transpose(1, 2).reshape(bs * heads, dim, int(h), int(w))
        g = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)
        out.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(bs * heads, 1, nq, L * P)
    o = (torch.stack(out, dim=-2).flatten(-2) * a).sum(-1).view(bs, heads * dim, nq)
    return o.transpose(1, 2).contiguous()


@pytest.mark.parametrize("bs,nq,heads,dim,shapes,P", [(2, 300, 8, 32, [(16, 20), (8, 10)], 4),
                                                     (1, 1000, 4, 16, [(25, 25)], 8),
                                                     (2, 50, 2, 48, [(7, 9), (5, 5), (3, 4)], 3)])
def test_forward_backward(cuda, bs, nq, heads, dim, shapes, P):
    torch.manual_seed(0)
    L = len(shapes)
    nk = sum(h * w for h, w in shapes)
    value = torch.randn(bs, nk, heads, dim, device=cuda)
    loc = torch.rand(bs, nq, heads, L, P, 2, device=cuda) * 1.3 - 0.15      # some samples out of bounds
    attn = torch.softmax(torch.randn(bs, nq, heads, L * P, device=cuda), -1).view(bs, nq, heads, L, P)
    sh = torch.tensor(shapes, dtype=torch.int64, device=cuda)
    st = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    v1, l1, a1 = (t.clone().requires_grad_(True) for t in (value, loc, attn))
    out = dbev.multi_scale_deformable_attn(v1, sh, st, l1, a1, 64)
    v2, l2, a2 = (t.double().clone().requires_grad_(True) for t in (value, loc, attn))
    ref = reference(v2, shapes, l2, a2)
    assert out.shape == ref.shape
    assert (out.double() - ref).abs().max() <= 1e-5 * ref.abs().max()
    w = torch.randn_like(out)
    (out * w).sum().backward()
    (ref * w.double()).sum().backward()
    for got, want in ((v1.grad, v2.grad), (l1.grad, l2.grad), (a1.grad, a2.grad)):
        assert (got.double() - want).abs().max() <= 1e-4 * want.abs().max()


def test_bevformer_size_runs(cuda):
    """BEVFormer temporal self attention scale: 200 x 200 queries, 8 heads x 32, 1 level, 4 points."""
    bs, nq, heads, dim = 2, 40000, 8, 32
    value = torch.randn(bs, 40000, heads, dim, device=cuda)
    loc = torch.rand(bs, nq, heads, 1, 4, 2, device=cuda)
    attn = torch.softmax(torch.randn(bs, nq, heads, 4, device=cuda), -1).view(bs, nq, heads, 1, 4)
    sh = torch.tensor([[200, 200]], dtype=torch.int64, device=cuda)
    out = dbev.multi_scale_deformable_attn(value, sh, sh.new_zeros(1), loc, attn)
    assert out.shape == (bs, nq, heads * dim) and torch.isfinite(out).all()


def test_cpu_raises():
    with pytest.raises(RuntimeError):
        dbev.multi_scale_deformable_attn(torch.zeros(1, 4, 1, 4), torch.tensor([[2, 2]]), torch.tensor([0]),
                                         torch.zeros(1, 3, 1, 1, 1, 2), torch.zeros(1, 3, 1, 1, 1))
