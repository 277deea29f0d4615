"""One table for SURVEY.md §8(d): every kernel group of the hot path at the BASELINE configs[1] /
configs[3] sizes -> CUDA-event time (median of 15, L2 flushed between launches), algorithmic
bytes / flops (the §8(d) formulas), achieved GB/s or TFLOP/s and the fraction of the measured peak
(MEASURED_PEAKS.json). Prints one JSON object.

    python tools/roofline_table.py > gpurun_out/roofline_table.json
"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import distill_bev_b200 as dbev
    from distill_bev_b200 import synthetic
    from distill_bev_b200.plugin.ops import spconv as sp
    dev = torch.device("cuda:0")
    peak, how = bench.measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def timed(fn, iters=15):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    def add(name, shape, fn, nbytes=None, flops=None, note=None):
        ms = timed(fn)
        r = {"op": name, "shape": shape, "ms": round(ms, 4)}
        if nbytes is not None:
            r["algorithmic_MB"] = round(nbytes / 1e6, 1)
            r["GBps"] = round(nbytes / ms / 1e6, 1)
            r["hbm_frac"] = round(nbytes / ms / 1e6 / peak, 3)
        if flops is not None:
            r["GFLOP"] = round(flops / 1e9, 2)
            r["TFLOPs"] = round(flops / ms / 1e9, 1)
        if note:
            r["note"] = note
        rows.append(r)

    hp = bench.HotPath(dev, 0)
    nf, C = bench.BATCH * bench.FRAMES, bench.C_TRANS
    geom = hp.vt.get_geometry(*hp.d_calib)
    n = geom.numel() // 3
    cams = nf * bench.N_CAMS
    add("lss_geometry (get_geometry)", "%d cams x %d frustum points" % (cams, n // cams),
        lambda: hp.vt.get_geometry(*hp.d_calib), nbytes=n * 12)
    plan = hp.vt.make_plan(geom, nf)          # carries point_cell (needed by the lift backward)
    kept = plan.num_kept()
    add("bev plan (keys + 3-pass radix sort + bounds + items)", "%d points -> %d kept" % (n, kept),
        lambda: hp.vt.make_plan(geom, nf), nbytes=n * 12 + 3 * n * 16 + n * 4,
        note="latency-bound multi-kernel sequence; bytes = geom read + 16 B/key/pass + order")
    x = torch.rand(n, C, device=dev)
    out = dbev.bev_pool_gather(x, plan)
    add("bev_pool gather forward (materialised rows)", "16 sample-frames, C=64, 128x128",
        lambda: dbev.bev_pool_gather(x, plan), nbytes=kept * C * 4 + kept * 4 + out.numel() * 4)
    xr = x.clone().requires_grad_(True)
    og = torch.rand_like(out)
    o2 = dbev.bev_pool_gather(xr, plan)

    def pool_bwd():
        o2.backward(og, retain_graph=True)
        xr.grad = None
    add("bev_pool gather backward", "same", pool_bwd, nbytes=out.numel() * 4 + n * 4 + n * C * 4)
    pix = cams * bench.FH * bench.FW
    add("fused lift+splat forward", "depth [%d,%d,16,44] x feat [%d,64,16,44] -> [16,64,128,128]" % (cams, bench.D, cams),
        lambda: dbev.lift_splat(hp.depth, hp.feat, plan), nbytes=(bench.D + C) * pix * 4 + kept * 4 + out.numel() * 4,
        note="algorithmic bytes exclude the never-materialised %d MB volume; the kernel is L2-bound (rows re-read from L2)"
             % (n * C * 4 // 1000000))
    cells = hp.vt.make_cells(geom, nf)
    add("bev_point_cells (geometry -> cell of every frustum point, no sort)", "%d points" % n,
        lambda: hp.vt.make_cells(geom, nf), nbytes=n * 16)
    add("sort-free lift+splat forward (float4 reductions into the L2-resident map, incl. its memset)", "same",
        lambda: dbev.lift_splat(hp.depth, hp.feat, cells), nbytes=(bench.D + C) * pix * 4 + n * 4 + 2 * out.numel() * 4,
        note="L2-bound: %d MB of 16-byte reductions stay in L2; HBM sees inputs + one write-back of the map" % (kept * C * 4 // 1000000))
    # teacher dense convs (tcgen05 TF32): one layer of each width at the B=8 sizes, and the whole stack
    from distill_bev_b200.plugin import dense_teacher as dt
    for hh, cc in ((256, 64), (128, 128), (64, 256)):
        xx = torch.randn(8, hh, hh, cc, device=dev)
        wp = torch.randn(cc, 9 * cc, device=dev) * 0.05
        sc, sh = torch.rand(cc, device=dev), torch.randn(cc, device=dev)
        oo = torch.empty_like(xx)
        add("conv3x3 + BN + ReLU (halo kernel, tcgen05 TF32)", "B=8, %d->%d, %dx%d" % (cc, cc, hh, hh),
            lambda xx=xx, wp=wp, sc=sc, sh=sh, oo=oo, cc=cc: dt.conv_nhwc(xx, wp, cc, 3, 3, 1, 1, sc, sh, relu=True, out=oo),
            nbytes=2 * xx.numel() * 4, flops=2.0 * xx.numel() * 9 * cc)
    with torch.no_grad():
        canvas = dbev.pillar_canvas(hp.d_points, hp.enc, hp.scat)
        add("SECOND + SECONDFPN (20 launches, PDL)", "canvas [8,64,512,512] -> [8,384,128,128]",
            lambda: hp.secfpn(hp.second(canvas)), flops=601.3e9)
    bev = dbev.lift_splat(hp.depth, hp.feat, plan)

    def lift_bwd():
        bev.backward(hp.bev_grad, retain_graph=True)
        hp.depth.grad = hp.feat.grad = None
    add("fused lift+splat backward", "same", lift_bwd,
        nbytes=out.numel() * 4 * 2 + (bench.D + C) * pix * 4 * 2 + n * 4)
    # teacher LiDAR path
    pts = torch.from_numpy(synthetic.make_lidar(1, 240000, seed=1)[0]).to(dev)
    vs, pcr = [0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    add("dynamic_voxelize", "240k points x 5", lambda: dbev.voxelization(pts, vs, pcr, -1, -1),
        nbytes=pts.shape[0] * (5 * 4 + 12))
    vox = dbev.Voxelization(vs, pcr, 10, (90000, 120000)).eval()
    v, c, k = vox(pts)
    add("hard_voxelize (deterministic, first-appearance order)", "240k points -> %d voxels x 10" % v.shape[0],
        lambda: vox(pts), nbytes=2 * pts.shape[0] * 20 + v.shape[0] * 16,
        note="includes the voxel_num read-back of the reference API; sort passes not counted in the bytes")
    coors = dbev.voxelization(pts, bench.PILLAR_VS, bench.PILLAR_RANGE, -1, -1)
    feats = torch.rand(pts.shape[0], 64, device=dev)
    ds = dbev.DynamicScatter(bench.PILLAR_VS, bench.PILLAR_RANGE, True)
    m = ds(feats, coors)[0].shape[0]
    add("dynamic scatter forward (mean)", "240k points x 64 ch -> %d pillars" % m, lambda: ds(feats, coors),
        nbytes=pts.shape[0] * (64 * 4 + 12) + m * (64 * 4 + 16) + pts.shape[0] * 4)
    with torch.no_grad():
        add("pillar encoder + scatter (dbev_pillar_canvas)", "8 x 30k points -> [8,64,512,512]",
            lambda: dbev.pillar_canvas(hp.d_points, hp.enc, hp.scat),
            nbytes=8 * 30000 * 20 + 8 * 64 * 512 * 512 * 4)
    # distillation head
    B, Cs, Ct, HW = bench.BATCH, bench.C_STUDENT, bench.C_TEACHER, bench.BEV * bench.BEV
    from distill_bev_b200.plugin.distill.adaptation import conv1x1
    with torch.no_grad():
        add("1x1 adaptation conv (tcgen05 TF32, incl. NCHW->channels-last transpose)", "B=8, 256->384, 128x128",
            lambda: conv1x1(hp.student, hp.adapt.weight, hp.adapt.bias),
            nbytes=B * HW * (Cs + Ct) * 4, flops=2.0 * B * HW * Cs * Ct)
    adapted = conv1x1(hp.student.detach(), hp.adapt.weight.detach(), hp.adapt.bias.detach()).requires_grad_(True)

    def loss():
        return dbev.fgd.fgd_distill_loss(hp.teacher, adapted, hp.boxes, bench.DISTILL_PARAMS, bench.TRAIN_CFG,
                                         spatial_adaptation=hp.spatial, heatmaps=hp.d_gt_hm,
                                         teacher_heatmaps=hp.teacher_logit, epoch=1)
    add("fgd distill loss forward (masks + 3 tensor reads)", "B=8, 384 ch, 128x128", loss, nbytes=3 * B * Ct * HW * 4)
    tot = sum(loss().values())

    def loss_bwd():
        tot.backward(retain_graph=True)
        adapted.grad = None
        hp.spatial.zero_grad(set_to_none=True)
    add("fgd distill loss backward", "same", loss_bwd, nbytes=3 * B * Ct * HW * 4)
    mask = (torch.rand(2, 1, 128, 128, device=dev) < 0.06).float()
    t2, s2 = hp.teacher[:2].contiguous(), adapted.detach()[:2].contiguous()
    K = int(mask.sum().item()) // 2
    add("affinity loss forward (select + gather + gram tiles)", "2 samples, ~%d cells each, 384 ch" % K,
        lambda: dbev.affinity.affinity_distill_loss(t2, s2, mask), flops=2 * 4.0 * K * K * Ct)
    # sparse teacher
    f = torch.rand(200000, 128, device=dev)
    lin = torch.randperm(2 * 2 * 200 * 200, device=dev)[:200000]
    cc = torch.stack([lin // 80000, (lin // 40000) % 2, (lin // 200) % 200, lin % 200], 1).int().contiguous()
    add("spconv dense() (SparseConvTensor.dense + view)", "200k voxels x 128 ch -> [2,256,200,200]",
        lambda: sp.dense_from_sparse(f, cc, [2, 200, 200], 2), nbytes=200000 * (512 + 16) + 2 * 256 * 200 * 200 * 4)
    print(json.dumps({"peak_GBps": peak, "peak_source": how, "timing": "CUDA events, median of 15, 256 MB L2 flush between launches, through the public plugin API (includes its host-side issue gaps)",
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
