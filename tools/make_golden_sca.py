"""Golden fixture for BEVFormer's spatial cross attention re-batching from the UNMODIFIED
forward bodies of SpatialCrossAttention (:76-174) and MSDeformableAttention3D (:273-399) of
mmdet3d/models/transformer_modules/spatial_cross_attention.py, cut out with `ast` (tools/ref_import.load_fgd_methods)
and executed on the CPU. mmcv is absent: ``multi_scale_deformable_attn_pytorch`` (third party, mmcv 1.6.0
ops/multi_scale_deform_attn.py) is restated from its published formula below - parity unpinned for that call only; the
re-batching, padding, slot accumulation and count division are the reference's own code.
Writes tests/golden/sca_small.npz. Build container only."""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402


def multi_scale_deformable_attn_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, num_heads, num_levels, num_points, _ = sampling_locations.shape
    value_list = value.split([int(H_ * W_) for H_, W_ in value_spatial_shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for level, (H_, W_) in enumerate(value_spatial_shapes):
        value_l_ = value_list[level].flatten(2).transpose(1, 2).reshape(bs * num_heads, embed_dims, int(H_), int(W_))
        sampling_grid_l_ = sampling_grids[:, :, :, level].transpose(1, 2).flatten(0, 1)
        sampling_value_list.append(F.grid_sample(value_l_, sampling_grid_l_, mode="bilinear", padding_mode="zeros",
                                                 align_corners=False))
    attention_weights = attention_weights.transpose(1, 2).reshape(bs * num_heads, 1, num_queries, num_levels * num_points)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) * attention_weights).sum(-1).view(
        bs, num_heads * embed_dims, num_queries)
    return output.transpose(1, 2).contiguous()


def main():
    rel = "mmdet3d/models/transformer_modules/spatial_cross_attention.py"
    ns = dict(multi_scale_deformable_attn_pytorch=multi_scale_deformable_attn_pytorch)
    sca_fns, _ = ref_import.load_fgd_methods(("forward",), cls_name="SpatialCrossAttention", relpath=rel, extra_ns=ns)
    msda_fns, _ = ref_import.load_fgd_methods(("forward",), cls_name="MSDeformableAttention3D", relpath=rel, extra_ns=ns)
    import distill_bev_b200  # noqa: F401  (CPU import: only the module classes are used, for their parameters)
    from distill_bev_b200.plugin.bevformer_attention import SpatialCrossAttention
    torch.manual_seed(0)
    C, cams, bs, nq, D = 64, 3, 2, 14 * 14, 4
    mod = SpatialCrossAttention(embed_dims=C, num_cams=cams, dropout=0.0,
                                deformable_attention=dict(type="MSDeformableAttention3D", embed_dims=C, num_heads=4,
                                                          num_levels=2, num_points=8))
    with torch.no_grad():                      # the reference initialises offsets / weights to structured constants
        mod.deformable_attention.sampling_offsets.weight.normal_(0, 0.05)
        mod.deformable_attention.attention_weights.weight.normal_(0, 0.2)
        mod.deformable_attention.attention_weights.bias.normal_(0, 0.2)
    shapes = torch.tensor([[8, 10], [4, 5]], dtype=torch.long)
    starts = torch.tensor([0, 80], dtype=torch.long)
    g = torch.Generator().manual_seed(1)
    query = torch.randn(bs, nq, C, generator=g).requires_grad_(True)
    query_pos = torch.randn(bs, nq, C, generator=g)
    value = torch.randn(cams, 100, bs, C, generator=g).requires_grad_(True)
    rpc = torch.rand(cams, bs, nq, D, 2, generator=g)
    # camera c sees a band of queries (different bands per batch element: the reference reads element 0's lists)
    bev_mask = torch.zeros(cams, bs, nq, D, dtype=torch.bool)
    for c in range(cams):
        for b in range(bs):
            lo = (c * 60 + b * 7) % nq
            sel = torch.arange(lo, lo + 90) % nq
            bev_mask[c, b, sel] = torch.rand(90, D, generator=g) > 0.4
    inner = types.SimpleNamespace(**{k: getattr(mod.deformable_attention, k) for k in (
        "batch_first", "num_heads", "num_levels", "num_points", "im2col_step", "value_proj", "sampling_offsets",
        "attention_weights")})
    deform = lambda **kw: msda_fns["forward"](inner, **kw)  # noqa: E731
    outer = types.SimpleNamespace(deformable_attention=deform, output_proj=mod.output_proj, dropout=mod.dropout,
                                  num_cams=cams, embed_dims=C)
    out = sca_fns["forward"](outer, query, value, value, query_pos=query_pos, reference_points_cam=rpc, bev_mask=bev_mask,
                             spatial_shapes=shapes, level_start_index=starts)
    go = torch.randn(out.shape, generator=g)
    out.backward(go)
    res = {"query": query.detach().numpy(), "query_pos": query_pos.numpy(), "value": value.detach().numpy(), "rpc": rpc.numpy(),
           "bev_mask": bev_mask.numpy(), "shapes": shapes.numpy(), "starts": starts.numpy(), "out": out.detach().numpy(),
           "go": go.numpy(), "d_query": query.grad.numpy(), "d_value": value.grad.numpy()}
    for k, v in mod.state_dict().items():
        res["sd/" + k] = v.numpy()
    for k, p in mod.named_parameters():
        res["grad/" + k] = p.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "sca_small.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6), "out", tuple(out.shape))


if __name__ == "__main__":
    main()
