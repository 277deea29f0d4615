"""How the gather-forward launch time depends on HOW it is timed (same kernel, same inputs):
back-to-back average, per-launch events, per-launch events with an L2 flush in between."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    import distill_bev_b200 as dbev
    from distill_bev_b200 import _lib, synthetic
    from distill_bev_b200.plugin.ops import bev_pool as bp
    nf = bench.BATCH * bench.FRAMES
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=8).to(dev)
    calib = [torch.from_numpy(a).to(dev) for a in synthetic.make_calibration(nf, bench.N_CAMS, seed=123)]
    geom = vt.get_geometry(*calib)
    plan = vt.make_plan(geom, nf, with_point_cell=False)
    n = geom.numel() // 3
    x = torch.rand(n, bench.C_TRANS, device=dev)
    shape, sB, sZ, sC = bp._out_strides(plan, bench.C_TRANS, "bz_c")
    out = torch.empty(shape, device=dev)
    lib = _lib.load()

    def launch():
        rc = lib.dbev_bev_pool_gather_forward(
            _lib.ptr(x), bench.C_TRANS, _lib.ptr(plan.order), _lib.ptr(plan.cell_start), _lib.ptr(plan.cell_end),
            _lib.ptr(plan.items), _lib.ptr(plan.n_items), plan.batch, plan.nz, plan.nslow, plan.nfast,
            sB, sZ, sC, _lib.ptr(out), _lib.stream_ptr(dev))
        _lib.check(rc, "gather")
    for _ in range(5):
        launch()
    res = {}
    for iters in (5, 30, 200):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            launch()
        b.record()
        torch.cuda.synchronize()
        res["back_to_back_%d" % iters] = a.elapsed_time(b) / iters
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, fl in (("per_launch", False), ("per_launch_flushed", True)):
        ts = []
        for _ in range(30):
            if fl:
                flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            launch()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name + "_med"] = statistics.median(ts)
        res[name + "_mean"] = sum(ts) / len(ts)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
