"""BASELINE.json configs[4]: bev_pool + distill-loss microbench sweep, D in {59, 118} (dbound step 1 / 0.5), BEV in
{128, 256, 512} (dx 0.8 / 0.4 / 0.2), C in {64, 256}, one GPU (per-sample ops: N GPUs = N replicas of these numbers).
Per point: the gather forward over materialised frustum rows (channels-last and NCHW output) as GB/s against the
measured HBM peak with the SURVEY §8(d) algorithmic bytes, the fused lift+splat forward + backward, and the fgd
distillation loss forward + backward (student = teacher = C_t channels at the BEV size, shipped recipe) as GB/s.
CUDA events, median of 9, L2 flushed between launches. Writes gpurun_out/configs4_sweep.json."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200 import _lib, synthetic  # noqa: E402
from distill_bev_b200.plugin.ops import bev_pool as bp  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak, how = bench.measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, iters=9):
        for _ in range(2):
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

    rows = []
    nf = 4                                  # sample-frames per launch (2 samples x 2 frames)
    lib = _lib.load()
    for D, dstep in ((59, 1.0), (118, 0.5)):
        for bev, cell in ((128, 0.8), (256, 0.4), (512, 0.2)):
            grid = dict(xbound=[-51.2, 51.2, cell], ybound=[-51.2, 51.2, cell], zbound=[-10.0, 10.0, 20.0],
                        dbound=[1.0, 60.0, dstep])
            vt = dbev.ViewTransformerLiftSplatShoot(grid_config=grid, numC_input=8).to(dev)
            assert vt.D == D
            calib = [torch.from_numpy(a).to(dev) for a in synthetic.make_calibration(nf, 6, seed=7)]
            geom = vt.get_geometry(*calib)
            n = geom.numel() // 3
            plan = vt.make_plan(geom, nf, with_point_cell=True)
            kept = plan.num_kept()
            t_plan = timed(lambda: vt.make_plan(geom, nf, with_point_cell=True))
            for C in (64, 256):
                x = torch.rand(n, C, device=dev)
                alg = kept * C * 4 + kept * 4 + nf * bev * bev * C * 4
                res = {"D": D, "bev": bev, "C": C, "sample_frames": nf, "rows_kept": int(kept), "rows": int(n),
                       "plan_ms": round(t_plan, 4), "gather_algorithmic_MB": round(alg / 1e6, 1)}
                for layout in ("cl", "bz_c"):
                    shape, sB, sZ, sC = bp._out_strides(plan, C, layout)
                    out = torch.empty(shape, device=dev)

                    def launch():
                        rc = lib.dbev_bev_pool_gather_forward(
                            _lib.ptr(x), C, _lib.ptr(plan.order), _lib.ptr(plan.cell_start), _lib.ptr(plan.cell_end),
                            _lib.ptr(plan.items), _lib.ptr(plan.n_items), plan.batch, plan.nz, plan.nslow, plan.nfast,
                            sB, sZ, sC, _lib.ptr(out), _lib.stream_ptr(dev))
                        _lib.check(rc, "gather")
                    ms = timed(launch)
                    tag = "channels_last" if layout == "cl" else "nchw"
                    res["gather_%s_ms" % tag] = round(ms, 4)
                    res["gather_%s_GBps" % tag] = round(alg / ms / 1e6, 1)
                    res["gather_%s_hbm_frac" % tag] = round(alg / ms / 1e6 / peak, 3)
                del x
                # fused lift + splat (never builds the rows): forward and backward
                depth = torch.randn(nf * 6, D, 16, 44, device=dev).softmax(1).requires_grad_(True)
                feat = torch.randn(nf * 6, C, 16, 44, device=dev).requires_grad_(True)
                og = torch.rand(nf, C, bev, bev, device=dev)

                def fused():
                    o = dbev.lift_splat(depth, feat, plan)
                    o.backward(og)
                    depth.grad = feat.grad = None
                res["lift_splat_fwd_bwd_ms (sorted plan)"] = round(timed(fused), 4)
                rows.append(res)
            del plan, geom
        # distillation loss at this BEV size is independent of D: measured once per BEV in the D = 59 pass
    loss_rows = []
    params = dict(bench.DISTILL_PARAMS, fp_as_foreground=["none"], fp_weight=0.0)
    for bev, cell in ((128, 0.8), (256, 0.4), (512, 0.2)):
        for C in (64, 256):
            B = 2
            tc = dict(grid_size=[1024, 1024, 40], point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], voxel_size=[0.1, 0.1, 0.2])
            boxes = [torch.from_numpy(b) for b, _ in synthetic.make_gt_boxes(B, seed=3)]
            teacher = torch.relu(torch.randn(B, C, bev, bev, device=dev))
            student = torch.relu(torch.randn(B, C, bev, bev, device=dev)).requires_grad_(True)
            spatial = torch.nn.Conv2d(1, 1, 3, padding=1).to(dev)

            def loss():
                l = dbev.fgd.fgd_distill_loss(teacher, student, boxes, params, tc, spatial_adaptation=spatial)
                sum(l.values()).backward()
                student.grad = None
                spatial.zero_grad(set_to_none=True)
            ms = timed(loss)
            alg = B * bev * bev * C * 4 * (2 + 3)      # forward: student + teacher once; backward: both again + the gradient
            loss_rows.append({"bev": bev, "C": C, "B": B, "fgd_loss_fwd_bwd_ms (incl. masks)": round(ms, 4),
                              "algorithmic_MB": round(alg / 1e6, 1), "GBps": round(alg / ms / 1e6, 1),
                              "hbm_frac": round(alg / ms / 1e6 / peak, 3)})
    res = {"hbm_peak_GBps": peak, "peak_source": how, "bev_pool": rows, "fgd_loss": loss_rows,
           "note": "one B200; per-sample ops: at N GPUs every rank runs a replica of these launches (no collective)"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs4_sweep.json"), "w") as f:
        json.dump(res, f, indent=1)
    for r in rows:
        print(r)
    for r in loss_rows:
        print(r)


if __name__ == "__main__":
    main()
