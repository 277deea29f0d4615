"""GPU parity: SpatialCrossAttention / MSDeformableAttention3D mirrors (plugin/bevformer_attention.py; per-camera
re-batching kernels csrc/sca_rebatch.cu + csrc/ms_deform_attn.cu) against tests/golden/sca_small.npz - the UNMODIFIED
forward bodies of the reference classes executed on the CPU (tools/make_golden_sca.py; mmcv's
multi_scale_deformable_attn_pytorch restated there from its published formula). Output 1e-4, gradients 1e-4 of max."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200  # noqa: F401
from distill_bev_b200.plugin import bevformer_attention as ba

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gd(golden_dir):
    return np.load(os.path.join(golden_dir, "sca_small.npz"))


def _module(gd, cuda):
    mod = ba.SpatialCrossAttention(embed_dims=64, num_cams=3, dropout=0.0,
                                   deformable_attention=dict(type="MSDeformableAttention3D", embed_dims=64, num_heads=4,
                                                             num_levels=2, num_points=8))
    mod.load_state_dict({k[3:]: torch.from_numpy(gd[k]) for k in gd.files if k.startswith("sd/")}, strict=True)
    return mod.to(cuda)


def test_camera_query_lists(gd, cuda):
    bev_mask = torch.from_numpy(gd["bev_mask"]).to(cuda)
    idx, pos, inv_count, max_len = ba.camera_query_lists(bev_mask)
    for cam in range(bev_mask.shape[0]):
        want = bev_mask[cam, 0].sum(-1).nonzero().squeeze(-1)              # spatial_cross_attention.py:131
        assert torch.equal(idx[cam, :len(want)].long(), want) and bool((idx[cam, len(want):] == -1).all())
        assert torch.equal(pos[cam, want].long(), torch.arange(len(want), device=cuda))
    assert max_len == max(int(bev_mask[c, 0].sum(-1).gt(0).sum()) for c in range(bev_mask.shape[0]))
    count = (bev_mask.sum(-1) > 0).permute(1, 2, 0).sum(-1).clamp(min=1.0)
    torch.testing.assert_close(inv_count, 1.0 / count)


def test_forward_backward_match_reference(gd, cuda):
    mod = _module(gd, cuda)
    query = torch.from_numpy(gd["query"]).to(cuda).requires_grad_(True)
    value = torch.from_numpy(gd["value"]).to(cuda).requires_grad_(True)
    out = mod(query, value, value, query_pos=torch.from_numpy(gd["query_pos"]).to(cuda),
              reference_points_cam=torch.from_numpy(gd["rpc"]).to(cuda), bev_mask=torch.from_numpy(gd["bev_mask"]).to(cuda),
              spatial_shapes=torch.from_numpy(gd["shapes"]).to(cuda), level_start_index=torch.from_numpy(gd["starts"]).to(cuda))
    want = torch.from_numpy(gd["out"]).to(cuda)
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-4)
    out.backward(torch.from_numpy(gd["go"]).to(cuda))
    for got, key in ((query.grad, "d_query"), (value.grad, "d_value")):
        w = torch.from_numpy(gd[key]).to(cuda)
        assert float((got - w).abs().max()) <= 1e-4 * float(w.abs().max()) + 1e-6, key
    for k, p in mod.named_parameters():
        w = torch.from_numpy(gd["grad/" + k]).to(cuda)
        assert float((p.grad - w).abs().max()) <= 2e-4 * float(w.abs().max()) + 1e-6, k


def test_rebatch_rows_are_deterministic_and_padding_is_zero(gd, cuda):
    bev_mask = torch.from_numpy(gd["bev_mask"]).to(cuda)
    idx, pos, inv_count, _ = ba.camera_query_lists(bev_mask)
    q = torch.randn(2, bev_mask.shape[2], 64, device=cuda, requires_grad=True)
    a = ba._Rebatch.apply(q, idx, pos)
    b = ba._Rebatch.apply(q, idx, pos)
    assert torch.equal(a, b)
    pad = (idx < 0)[None, :, :, None].expand_as(a)
    assert float(a[pad].abs().max()) == 0.0
    s1 = ba._Slots.apply(a, idx, pos, inv_count)
    s2 = ba._Slots.apply(a, idx, pos, inv_count)
    assert torch.equal(s1, s2)
    # a query seen by k cameras gets k copies / count = itself back (count from each element's own mask, lists from element 0)
    hit0 = (bev_mask[:, 0].sum(-1) > 0).sum(0).float()
    torch.testing.assert_close(s1, q.detach() * (hit0[None, :, None] * inv_count[:, :, None]), rtol=1e-6, atol=1e-6)
