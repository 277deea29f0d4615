"""GPU parity: bev_pool / voxel_pooling CUDA path (through the C-ABI shims) vs
the CPU oracle and the committed reference fixtures.

Tolerances: indices / kept sets / empty cells are exact; pooled features are
fp32 sums compared with the fp64-accumulated oracle at rtol 1e-5 (north_star
bound is 1e-3 relative; the reference's own cumsum path is only good to ~1e-3).
"""
import json
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev
from distill_bev_b200 import synthetic
from oracle import lss_oracle

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-5


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_voxel_pooling_golden_small(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "lss_small.npz"))
    x = _t(g["x"], cuda).requires_grad_(True)
    out = dbev.voxel_pooling(_t(g["geom"], cuda), x, g["bx"], g["dx"], g["nx"])
    assert tuple(out.shape) == g["out_cumsum"].shape
    o = out.detach().cpu().numpy()
    np.testing.assert_allclose(o, g["out_accelerated"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(o, g["out_cumsum"], rtol=1e-3, atol=1e-3)
    np.testing.assert_array_equal(o == 0, g["out_accelerated"] == 0)
    (out * _t(g["out_weight"], cuda)).sum().backward()
    np.testing.assert_array_equal(x.grad.cpu().numpy(), g["x_grad"])


def test_voxel_pooling_golden_edges(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "lss_edge.npz"))
    x = _t(g["x"], cuda).requires_grad_(True)
    geom = _t(g["geom"], cuda)
    plan = dbev.bev_plan_from_geom(geom, 3, g["bx"], g["dx"], g["nx"])
    assert plan.num_kept() == 10
    out = dbev.voxel_pooling(geom, x, g["bx"], g["dx"], g["nx"], plan=plan)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out_cumsum"], rtol=RTOL, atol=1e-6)
    (out * _t(g["out_weight"], cuda)).sum().backward()
    np.testing.assert_array_equal(x.grad.cpu().numpy(), g["x_grad"])


def test_bev_pool_golden_quickcumsum(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "quickcumsum.npz"))
    B, D, H, W = int(g["B"]), int(g["D"]), int(g["H"]), int(g["W"])
    feats = _t(g["feats"], cuda).requires_grad_(True)
    out = dbev.bev_pool(feats, _t(g["coords"], cuda), B, D, H, W)
    assert tuple(out.shape) == g["dense"].shape
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["dense"], rtol=1e-4, atol=1e-4)
    # reference-ABI path: pre-sorted rows + ranks (QuickCumsumCuda contract)
    order = torch.from_numpy(g["sort_index"]).to(cuda)
    xs = _t(g["feats"], cuda)[order].requires_grad_(True)
    cs = _t(g["coords"], cuda)[order]
    ranks = cs[:, 0] * (W * D * B) + cs[:, 1] * (D * B) + cs[:, 2] * B + cs[:, 3]
    dense = dbev.QuickCumsumCuda.apply(xs, cs, ranks, B, D, H, W)   # [B, D, H, W, C]
    np.testing.assert_allclose(dense.permute(0, 4, 1, 2, 3).detach().cpu().numpy(), g["dense"],
                               rtol=1e-4, atol=1e-4)
    og = torch.zeros_like(dense)
    gp = torch.from_numpy(g["geom_pooled"]).to(cuda)
    og[gp[:, 3], gp[:, 2], gp[:, 0], gp[:, 1]] = _t(g["weight"], cuda)
    dense.backward(og)
    np.testing.assert_array_equal(xs.grad.cpu().numpy(), g["x_sorted_grad"])


@pytest.mark.parametrize("C", [4, 8, 16, 32, 64, 80, 128, 256, 6, 33])
def test_bev_pool_channel_widths(cuda, C):
    rng = np.random.RandomState(C)
    B, D, H, W, n = 2, 2, 40, 37, 20000
    coords = np.stack([rng.randint(0, H, n), rng.randint(0, W, n), rng.randint(0, D, n),
                       rng.randint(0, B, n)], 1).astype(np.int64)
    coords[: n // 4, :2] = coords[0, :2]  # one very long interval
    feats = rng.random_sample((n, C)).astype(np.float32)
    ft = _t(feats, cuda).requires_grad_(True)
    out = dbev.bev_pool(ft, _t(coords, cuda), B, D, H, W)
    ref = lss_oracle.bev_pool(feats, coords, B, D, H, W)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=RTOL, atol=1e-3 * RTOL * n)
    w = rng.random_sample(ref.shape).astype(np.float32)
    (out * _t(w, cuda)).sum().backward()
    np.testing.assert_array_equal(ft.grad.cpu().numpy(), lss_oracle.bev_pool_backward(w, coords))
    # int32 coords take the same path
    out32 = dbev.bev_pool(ft.detach(), _t(coords.astype(np.int32), cuda), B, D, H, W)
    assert torch.equal(out32, out.detach())


@pytest.mark.parametrize("rpi", [1, 7, 64, 1000000])
def test_work_item_split_is_invisible(cuda, rpi):
    """The work-item split (rows_per_item) never changes what is computed."""
    rng = np.random.RandomState(3)
    B, D, H, W, n, C = 1, 1, 48, 70, 60000, 64
    coords = np.stack([rng.randint(0, H, n), rng.randint(0, W, n), rng.randint(0, D, n),
                       rng.randint(0, B, n)], 1).astype(np.int64)
    coords[: n // 3] = coords[0]            # one cell holds a third of all rows
    coords[n // 3: n // 2, 1] = 5           # one heavy column
    feats = rng.random_sample((n, C)).astype(np.float32)
    ft, ct = _t(feats, cuda), _t(coords, cuda)
    plan = dbev.bev_plan_from_coords(ct, B, D, H, W, rows_per_item=rpi)
    out = dbev.bev_pool_gather(ft.requires_grad_(True), plan, layout="b_c_z")
    ref = lss_oracle.bev_pool(feats, coords, B, D, H, W)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=RTOL, atol=1e-2)
    w = rng.random_sample(ref.shape).astype(np.float32)
    (out * _t(w, cuda)).sum().backward()
    np.testing.assert_array_equal(ft.grad.cpu().numpy(), lss_oracle.bev_pool_backward(w, coords))
    # same plan parameters => bit-identical reruns (fixed in-cell order, fixed fix-up order);
    # a different split only regroups fp32 partial sums
    again = dbev.bev_pool_gather(ft.detach(), dbev.bev_plan_from_coords(ct, B, D, H, W, rows_per_item=rpi),
                                 layout="b_c_z")
    assert torch.equal(out.detach(), again)
    base = dbev.bev_pool(ft.detach(), ct, B, D, H, W)
    torch.testing.assert_close(out.detach(), base, rtol=1e-5, atol=1e-2)


def test_bev_pool_empty_and_out_of_range(cuda):
    out = dbev.bev_pool(torch.zeros(0, 16, device=cuda), torch.zeros(0, 4, dtype=torch.long, device=cuda),
                        2, 1, 8, 8)
    assert tuple(out.shape) == (2, 16, 1, 8, 8) and float(out.abs().sum()) == 0.0
    coords = torch.tensor([[0, 0, 0, 0], [8, 0, 0, 0], [-1, 3, 0, 1], [7, 7, 0, 1]], device=cuda)
    feats = torch.ones(4, 16, device=cuda)
    out = dbev.bev_pool(feats, coords, 2, 1, 8, 8)
    assert float(out.sum()) == 32.0  # the two out-of-grid rows are dropped, not written out of bounds
    assert float(out[0, :, 0, 0, 0].sum()) == 16.0 and float(out[1, :, 0, 7, 7].sum()) == 16.0


def test_reference_abi_interval_kernels(cuda):
    rng = np.random.RandomState(11)
    b, d, h, w, c, n = 2, 1, 16, 16, 64, 5000
    coords = np.stack([rng.randint(0, h, n), rng.randint(0, w, n), rng.randint(0, d, n),
                       rng.randint(0, b, n)], 1).astype(np.int64)
    feats = rng.random_sample((n, c)).astype(np.float32)
    order, _, starts, lengths = lss_oracle.sorted_intervals(coords, b, d, h, w)
    xs, cs = feats[order], coords[order].astype(np.int32)
    ext = dbev.bev_pool_ext
    out = ext.bev_pool_forward(_t(xs, cuda), _t(cs, cuda), _t(lengths, cuda), _t(starts, cuda), b, d, h, w)
    ref = lss_oracle.bev_pool_interval_forward(xs, cs, starts, lengths, b, d, h, w)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    og = rng.random_sample(ref.shape).astype(np.float32)
    xg = ext.bev_pool_backward(_t(og, cuda), _t(cs, cuda), _t(lengths, cuda), _t(starts, cuda), b, d, h, w)
    np.testing.assert_array_equal(xg.cpu().numpy(),
                                  lss_oracle.bev_pool_interval_backward(og, cs, starts, lengths, n))


def _config1_inputs(batch, seed=0, C=64, bev=128, dstep=1.0):
    grid = synthetic.grid_config(bev, dstep)
    dx, bx, nx = lss_oracle.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    frustum = lss_oracle.create_frustum(synthetic.NUSC_INPUT_SIZE, 16, grid["dbound"])
    calib = synthetic.make_calibration(batch, 6, seed=seed)
    geom = lss_oracle.get_geometry(frustum, *calib)
    D, fH, fW = frustum.shape[:3]
    x = synthetic.make_frustum_feats(batch * 6 * D * fH * fW, C, seed=seed).reshape(batch, 6, D, fH, fW, C)
    return geom, x, bx, dx, nx


def test_voxel_pooling_config1_full_size_vs_oracle(cuda):
    """BASELINE.json configs[0]: 1 sample, 6 cams, D=59, 16x44, C=64 -> 128x128."""
    geom, x, bx, dx, nx = _config1_inputs(1)
    xt = _t(x, cuda).requires_grad_(True)
    gt = _t(geom, cuda)
    plan = dbev.bev_plan_from_geom(gt, 1, bx, dx, nx)
    idx, kept = lss_oracle.voxel_indices(geom, bx, dx, nx)
    assert plan.num_kept() == int(kept.sum())                      # exact kept set size
    out = dbev.voxel_pooling(gt, xt, bx, dx, nx, plan=plan)
    ref = lss_oracle.voxel_pooling(geom, x, bx, dx, nx)
    o = out.detach().cpu().numpy()
    assert o.shape == (1, 64, 128, 128)
    np.testing.assert_array_equal(o == 0, ref == 0)               # same set of empty cells
    np.testing.assert_allclose(o, ref, rtol=RTOL, atol=ATOL)
    # plan internals are exact integers: order is the stable sort of the cell keys
    n_i = nx.astype(np.int64)
    key = np.where(kept, (idx[:, 2] * n_i[1] + idx[:, 1]) * n_i[0] + idx[:, 0], n_i.prod())
    np.testing.assert_array_equal(plan.order.cpu().numpy()[: key.size], np.argsort(key, kind="stable"))
    w = np.random.RandomState(1).random_sample(ref.shape).astype(np.float32)
    (out * _t(w, cuda)).sum().backward()
    xg = lss_oracle.voxel_pooling_backward(geom, w, 64, bx, dx, nx)
    np.testing.assert_array_equal(xt.grad.cpu().numpy().reshape(-1, 64), xg)


def test_voxel_pooling_batch8_properties(cuda):
    """configs[1] batch (B=8): size-independent properties at full size.

    linearity: pool(a*x + y) = a*pool(x) + pool(y); conservation: the sum over
    the BEV map equals the sum over kept rows; determinism: bit-identical reruns;
    batch independence: sample b of the batched call equals the single call.
    """
    B = 8
    geom, x, bx, dx, nx = _config1_inputs(B, seed=5)
    gt, xt = _t(geom, cuda), _t(x, cuda)
    plan = dbev.bev_plan_from_geom(gt, B, bx, dx, nx)
    out = dbev.voxel_pooling(gt, xt, bx, dx, nx, plan=plan)
    out2 = dbev.voxel_pooling(gt, xt, bx, dx, nx)               # fresh plan
    assert torch.equal(out, out2)
    y = torch.rand_like(xt)
    lhs = dbev.voxel_pooling(gt, 0.5 * xt + y, bx, dx, nx, plan=plan)
    rhs = 0.5 * out + dbev.voxel_pooling(gt, y, bx, dx, nx, plan=plan)
    torch.testing.assert_close(lhs, rhs, rtol=1e-4, atol=1e-3)
    _, kept = lss_oracle.voxel_indices(geom, bx, dx, nx)
    kept_t = torch.from_numpy(kept).to(cuda)
    total = xt.reshape(-1, 64)[kept_t].double().sum()
    assert abs(float(out.double().sum()) - float(total)) / float(total) < 1e-6
    single = dbev.voxel_pooling(gt[3:4].contiguous(), xt[3:4].contiguous(), bx, dx, nx)
    assert torch.equal(single[0], out[3])


@pytest.mark.parametrize("bev,dstep,C", [(256, 1.0, 64), (128, 0.5, 64), (512, 1.0, 64), (128, 1.0, 256)])
def test_voxel_pooling_sweep_vs_oracle(cuda, bev, dstep, C):
    """BASELINE.json configs[4] sweep points (one sample each)."""
    geom, x, bx, dx, nx = _config1_inputs(1, seed=2, C=C, bev=bev, dstep=dstep)
    out = dbev.voxel_pooling(_t(geom, cuda), _t(x, cuda), bx, dx, nx)
    ref = lss_oracle.voxel_pooling(geom, x, bx, dx, nx)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=RTOL, atol=ATOL)


def test_lift_splat_golden_small(cuda, golden_dir):
    """Fused lift+splat vs the reference's own lift (outer product, permute) + voxel_pooling."""
    g = np.load(os.path.join(golden_dir, "lss_small.npz"))
    B = g["rots"].shape[0]
    plan = dbev.bev_plan_from_geom(_t(g["geom"], cuda), B, g["bx"], g["dx"], g["nx"], with_point_cell=True)
    depth = _t(g["lift_depth"], cuda).requires_grad_(True)
    feat = _t(g["lift_feat"], cuda).requires_grad_(True)
    out = dbev.lift_splat(depth, feat, plan)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["lift_out"], rtol=RTOL, atol=ATOL)
    (out * _t(g["out_weight"], cuda)).sum().backward()
    np.testing.assert_allclose(depth.grad.cpu().numpy(), g["lift_ddepth"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(feat.grad.cpu().numpy(), g["lift_dfeat"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("C,frames", [(64, 2), (80, 1), (256, 1)])
def test_lift_splat_full_size_equals_materialised_path(cuda, C, frames):
    """configs[0]/[1] frustum (6 cams, D=59, 16x44): fused result == voxel_pooling of the
    materialised volume (same kernels, same summation order -> tight), grads vs oracle."""
    geom, _, bx, dx, nx = _config1_inputs(frames, seed=4, C=4)
    rng = np.random.RandomState(C)
    BN, D, fH, fW = frames * 6, 59, 16, 44
    dl = rng.randn(BN, D, fH, fW).astype(np.float32)
    depth = np.exp(dl) / np.exp(dl).sum(1, keepdims=True)
    feat = rng.randn(BN, C, fH, fW).astype(np.float32)
    gt = _t(geom, cuda)
    plan = dbev.bev_plan_from_geom(gt, frames, bx, dx, nx, with_point_cell=True)
    dt, ft = _t(depth, cuda).requires_grad_(True), _t(feat, cuda).requires_grad_(True)
    out = dbev.lift_splat(dt, ft, plan)
    vol = (dt.detach().unsqueeze(1) * ft.detach().unsqueeze(2)).view(frames, 6, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    ref = dbev.voxel_pooling(gt, vol, bx, dx, nx, plan=plan)
    torch.testing.assert_close(out.detach(), ref, rtol=1e-5, atol=1e-5)
    w = rng.random_sample(tuple(out.shape)).astype(np.float32)
    (out * _t(w, cuda)).sum().backward()
    if C == 64:
        dd, df = lss_oracle.lift_splat_backward(geom, depth, feat, w, frames, 6, bx, dx, nx)
        np.testing.assert_allclose(dt.grad.cpu().numpy(), dd, rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(ft.grad.cpu().numpy(), df, rtol=1e-3, atol=1e-4)
    else:
        vol2 = vol.clone().requires_grad_(True)
        (dbev.voxel_pooling(gt, vol2, bx, dx, nx, plan=plan) * _t(w, cuda)).sum().backward()
        gx = vol2.grad.reshape(BN, D, fH, fW, C)
        torch.testing.assert_close(dt.grad, torch.einsum("bdhwc,bchw->bdhw", gx, ft.detach()), rtol=1e-3, atol=1e-3)
        torch.testing.assert_close(ft.grad, torch.einsum("bdhwc,bdhw->bchw", gx, dt.detach()), rtol=1e-3, atol=1e-3)


def test_transpose_batched(cuda):
    x = torch.rand(3, 70, 45, device=cuda)
    assert torch.equal(dbev.transpose_batched(x, 3, 70, 45), x.transpose(1, 2).contiguous())


def test_sort_free_lift_splat_matches_sorted(cuda):
    """Opt-in sort-free lift+splat (vector float reductions, no plan) == the plan-based kernel up to
    fp32 summation order, forward and gradients; output is channels_last."""
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=8).to(cuda)
    nf = 4
    calib = [torch.from_numpy(a).to(cuda) for a in synthetic.make_calibration(nf, 6, seed=5)]
    geom = vt.get_geometry(*calib)
    torch.manual_seed(0)
    depth = torch.randn(nf * 6, 59, 16, 44, device=cuda).softmax(1)
    feat = torch.randn(nf * 6, 64, 16, 44, device=cuda)
    og = torch.rand(nf, 64, 128, 128, device=cuda)
    outs = []
    for mk in (lambda: vt.make_plan(geom, nf), lambda: vt.make_cells(geom, nf)):
        d, f = depth.clone().requires_grad_(True), feat.clone().requires_grad_(True)
        out = dbev.lift_splat(d, f, mk())
        out.backward(og)
        outs.append((out.detach(), d.grad, f.grad))
    assert outs[1][0].is_contiguous(memory_format=torch.channels_last)
    for a, b in zip(outs[0], outs[1]):
        torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-5 * float(a.abs().max()))
    pc = vt.make_cells(geom, nf).point_cell
    assert torch.equal(pc, vt.make_plan(geom, nf).point_cell)


def test_sort_free_lift_splat_frames_is_the_channel_concat(cuda):
    """frames=2 (BEVDepth4D): with the frame index numbered last in the cell id the splat's channels-last map IS
    torch.cat([frame 0, frame 1], dim=1) of the per-sample-frame result (bevdet.py:300-320); same for the gradients."""
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=8).to(cuda)
    nf, frames = 6, 2
    calib = [torch.from_numpy(a).to(cuda) for a in synthetic.make_calibration(nf, 6, seed=9)]
    geom = vt.get_geometry(*calib)
    torch.manual_seed(1)
    depth = torch.randn(nf * 6, 59, 16, 44, device=cuda).softmax(1)
    feat = torch.randn(nf * 6, 64, 16, 44, device=cuda)
    og = torch.rand(nf // frames, frames * 64, 128, 128, device=cuda).contiguous(memory_format=torch.channels_last)
    d0, f0 = depth.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    ref = dbev.lift_splat(d0, f0, vt.make_cells(geom, nf))                        # [nf, 64, 128, 128]
    ref_cat = torch.cat([ref[0::2], ref[1::2]], 1)
    ref_cat.backward(og)
    d1, f1 = depth.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    out = dbev.lift_splat(d1, f1, vt.make_cells(geom, nf, frames=frames))
    assert tuple(out.shape) == (nf // frames, frames * 64, 128, 128)
    assert out.is_contiguous(memory_format=torch.channels_last)
    out.backward(og)
    torch.testing.assert_close(out.detach(), ref_cat.detach(), rtol=1e-4, atol=1e-5 * float(ref_cat.abs().max()))
    torch.testing.assert_close(d1.grad, d0.grad, rtol=1e-4, atol=1e-5 * float(d0.grad.abs().max()))
    torch.testing.assert_close(f1.grad, f0.grad, rtol=1e-4, atol=1e-5 * float(f0.grad.abs().max()))
    # a gradient that is not channels_last goes through the transpose path
    d2, f2 = depth.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    dbev.lift_splat(d2, f2, vt.make_cells(geom, nf, frames=frames)).backward(og.contiguous())
    torch.testing.assert_close(d2.grad, d1.grad, rtol=1e-5, atol=1e-6 * float(d1.grad.abs().max()))
    with pytest.raises(RuntimeError):
        vt.make_cells(geom, nf, frames=4)           # 6 sample-frames are not a multiple of 4


@pytest.mark.parametrize("C", [64, 32, 80, 128])
def test_channels_last_gather_is_bit_identical_to_the_nchw_layout(cuda, C):
    """layout 'cl' / voxel_pooling(channels_last=True): finished cells are stored as whole rows straight from the
    accumulator (no transposing epilogue); same accumulation order -> bit-identical values, forward and backward,
    including empty cells (zero rows) and cells split between the two workers of a warp."""
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=8).to(cuda)
    nf = 3
    calib = [torch.from_numpy(a).to(cuda) for a in synthetic.make_calibration(nf, 6, seed=11)]
    geom = vt.get_geometry(*calib)
    torch.manual_seed(C)
    x = torch.rand(nf, 6, 59, 16, 44, C, device=cuda)
    for with_pc in (True, False):
        plan = vt.make_plan(geom, nf, with_point_cell=with_pc)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ref = vt.voxel_pooling(geom, xa, plan=plan)
        got = vt.voxel_pooling(geom, xb, plan=plan, channels_last=True)
        assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(got, ref)
        g = torch.rand_like(ref)
        ref.backward(g)
        got.backward(g.contiguous(memory_format=torch.channels_last))
        assert torch.equal(xb.grad, xa.grad)
