"""The reference's own CUDA kernels ("the kernels to beat") timed on the same B200 beside ours, same
inputs, CUDA events (median of 7 after 3 warm-ups):

  mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu:20-98      bev_pool_kernel / bev_pool_grad_kernel (+ bev_pool.py's prelude)
  mmdet3d/ops/voxel/src/voxelization_cuda.cu:231-528   dynamic / hard voxelize (deterministic and not)
  mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-308 dynamic_point_to_voxel_forward

compiled UNMODIFIED for sm_100 by oracle/build_oracle.py into oracle/_ref/ (evidence only; nothing in the product
imports them). Writes gpurun_out/reference_cuda_kernels.json (copied to profiles/r02_reference_cuda_kernels.json).
"""
import glob
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200 import synthetic  # noqa: E402


def load_ref(name):
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + "*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location(name, so[0])
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    return ext


def timed(fn, iters=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2], out


def bev_pool_section(dev, res):
    ext = load_ref("ref_bev_pool_ext")
    B, NF, C, D, H, W = 8, 16, 64, 1, 128, 128
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=8).to(dev)
    calib = [torch.from_numpy(a).to(dev) for a in synthetic.make_calibration(NF, 6, seed=1000)]
    geom = vt.get_geometry(*calib)                                   # [NF, 6, 59, 16, 44, 3]
    n = geom.numel() // 3
    # voxel_pooling's index math (view_transformer_mine.py:150-161), then bev_pool.py:83-97's prelude
    dx, bx, nx = vt.dx, vt.bx, vt.nx
    idx = ((geom - (bx - dx / 2.0)) / dx).long().view(n, 3)
    batch_ix = torch.arange(NF, device=dev).view(NF, 1).expand(NF, n // NF).reshape(n, 1)
    idx = torch.cat([idx, batch_ix], 1)
    kept = (idx[:, 0] >= 0) & (idx[:, 0] < int(nx[0])) & (idx[:, 1] >= 0) & (idx[:, 1] < int(nx[1])) & (idx[:, 2] >= 0) & (
        idx[:, 2] < int(nx[2]))
    feats_all = torch.rand(n, C, device=dev)
    feats, coords = feats_all[kept].contiguous(), idx[kept].contiguous()

    def ref_op():
        """bev_pool.py:83-97 + QuickCumsumCuda.forward :38-60, restated with the same torch ops."""
        ranks = coords[:, 0] * (W * D * NF) + coords[:, 1] * (D * NF) + coords[:, 2] * NF + coords[:, 3]
        indices = ranks.argsort()
        x, g, r = feats[indices], coords[indices], ranks[indices]
        k = torch.ones(x.shape[0], device=dev, dtype=torch.bool)
        k[1:] = r[1:] != r[:-1]
        starts = torch.where(k)[0].int()
        lengths = torch.zeros_like(starts)
        lengths[:-1] = starts[1:] - starts[:-1]
        lengths[-1] = x.shape[0] - starts[-1]
        g = g.int()
        out = ext.bev_pool_forward(x, g, lengths, starts, NF, D, H, W)
        return out.permute(0, 4, 1, 2, 3).contiguous(), (x, g, lengths, starts)

    t_ref_op, (out_ref, (xs, gs, lens, starts)) = timed(ref_op)
    t_ref_kernel, _ = timed(lambda: ext.bev_pool_forward(xs, gs, lens, starts, NF, D, H, W))
    grad = torch.rand(NF, D, H, W, C, device=dev)
    t_ref_bwd, _ = timed(lambda: ext.bev_pool_backward(grad, gs, lens, starts, NF, D, H, W))
    t_our_op, out_ours = timed(lambda: dbev.bev_pool(feats, coords, NF, D, H, W))
    plan = dbev.bev_plan_from_coords(coords, NF, D, H, W, fast_axis=1)
    t_our_kernel, _ = timed(lambda: dbev.bev_pool_gather(feats, plan, layout="b_c_z"))
    t_our_abi, out_abi = timed(lambda: dbev.bev_pool_ext.bev_pool_forward(xs, gs, lens, starts, NF, D, H, W))
    err = float((out_ours - out_ref).abs().max() / out_ref.abs().max())
    err_abi = float((out_abi - ext.bev_pool_forward(xs, gs, lens, starts, NF, D, H, W)).abs().max())
    rows = int(feats.shape[0])
    res["bev_pool"] = {
        "shape": "16 sample-frames, %d rows kept of %d, C=64, 128x128 (configs[1])" % (rows, n),
        "reference_op_ms (argsort + gathers + where + bev_pool_kernel + permute)": round(t_ref_op, 4),
        "reference_kernel_ms (torch::zeros + bev_pool_kernel)": round(t_ref_kernel, 4),
        "reference_grad_kernel_ms": round(t_ref_bwd, 4),
        "ours_op_ms (plan + gather, same API)": round(t_our_op, 4), "ours_gather_kernel_ms (plan cached)": round(t_our_kernel, 4),
        "ours_reference_abi_interval_kernel_ms": round(t_our_abi, 4),
        "max_rel_diff_ours_vs_reference": err, "max_abs_diff_interval_abi": err_abi}


def voxel_section(dev, res):
    ext = load_ref("ref_voxel_layer_cuda")
    out = {}
    for name, n_pts, vs, rng, max_pts, max_vox in (
            ("pillar 30k", 30000, [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 20, 30000),
            ("pillar 240k", 240000, [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 20, 30000),
            ("sparse 240k", 240000, [0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 10, 90000)):
        pts = torch.from_numpy(synthetic.make_lidar(1, n_pts, seed=5)[0]).to(dev)
        row = {}
        co_r = torch.zeros(n_pts, 3, dtype=torch.int32, device=dev)
        co_o = torch.zeros(n_pts, 3, dtype=torch.int32, device=dev)
        row["dynamic_voxelize ref_ms"] = round(timed(lambda: ext.dynamic_voxelize(pts, co_r, vs, rng, 3))[0], 4)
        row["dynamic_voxelize ours_ms"] = round(timed(lambda: dbev.voxel_layer.dynamic_voxelize(pts, co_o, vs, rng, 3))[0], 4)
        row["dynamic coors equal"] = bool(torch.equal(co_r, co_o))

        def hard(mod, det):
            v = torch.zeros(max_vox, max_pts, pts.shape[1], device=dev)
            c = torch.zeros(max_vox, 3, dtype=torch.int32, device=dev)
            k = torch.zeros(max_vox, dtype=torch.int32, device=dev)
            m = mod.hard_voxelize(pts, v, c, k, vs, rng, max_pts, max_vox, 3, det)
            return v, c, k, m
        iters = 3 if n_pts > 100000 else 7
        t, (v_r, c_r, k_r, m_r) = timed(lambda: hard(ext, True), iters=iters, warm=1)
        row["hard_voxelize deterministic ref_ms"] = round(t, 3)
        t, _ = timed(lambda: hard(ext, False), iters=iters, warm=1)
        row["hard_voxelize non-deterministic ref_ms"] = round(t, 3)
        t, (v_o, c_o, k_o, m_o) = timed(lambda: hard(dbev.voxel_layer, True), iters=iters, warm=1)
        row["hard_voxelize ours_ms (deterministic order)"] = round(t, 3)
        row["hard voxel count ref / ours"] = [int(m_r), int(m_o)]
        row["hard outputs equal"] = bool(int(m_r) == int(m_o) and torch.equal(c_r[:m_r], c_o[:m_o]) and
                                         torch.equal(k_r[:m_r], k_o[:m_o]) and torch.equal(v_r[:m_r], v_o[:m_o]))
        feats = torch.rand(n_pts, 64, device=dev)
        co = co_o.clone()
        row["dynamic_scatter(max, C=64) ref_ms"] = round(timed(lambda: ext.dynamic_point_to_voxel_forward(feats, co, "max"))[0], 4)
        row["dynamic_scatter(max, C=64) ours_ms"] = round(timed(
            lambda: dbev.voxel_layer.dynamic_point_to_voxel_forward(feats, co, "max"))[0], 4)
        out[name] = row
    res["voxel"] = out


def main():
    dev = torch.device("cuda:0")
    res = {"note": "reference kernels = the unmodified .cu/.cpp of /root/reference compiled for sm_100 (oracle/build_oracle.py)"}
    for name, fn in (("bev_pool", bev_pool_section), ("voxel", voxel_section)):
        try:
            fn(dev, res)
        except Exception as exc:  # noqa: BLE001
            res[name + "_error"] = repr(exc)[:400]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "reference_cuda_kernels.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
