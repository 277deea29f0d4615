"""Where does the bench step's time go? CPU issue time vs GPU time, GPU busy (sum of kernel
durations from CUPTI) vs step span, and the top kernels by in-step (warm) duration."""
import collections
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402


def phases(hp):
    """CPU issue time (queue empty at phase start) and GPU time of every phase of the step."""
    import distill_bev_b200 as dbev
    res = collections.OrderedDict()

    def timed(name, fn, reps=10):
        cpu = gpu = 0.0
        val = None
        for _ in range(reps):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            t0 = time.perf_counter()
            a.record()
            val = fn()
            b.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            cpu += (t1 - t0) * 1e3
            gpu += a.elapsed_time(b)
        res[name] = {"cpu_issue_ms": round(cpu / reps, 4), "gpu_ms": round(gpu / reps, 4)}
        return val

    nf = bench.BATCH * bench.FRAMES
    geom = timed("A.geometry", lambda: hp.vt.get_geometry(*hp.d_calib))
    plan = timed("A.plan", lambda: hp.vt.make_plan(geom, nf))
    state = {}

    def fwd():
        state["bev"] = dbev.lift_splat(hp.depth, hp.feat, plan)

    def bwd():
        state["bev"].backward(hp.bev_grad, retain_graph=True)
        hp.depth.grad = hp.feat.grad = None
    timed("A.lift_fwd", fwd)
    timed("A.lift_bwd", bwd)

    def pillars():
        with torch.no_grad():
            return dbev.pillar_canvas(hp.d_points, hp.enc, hp.scat)
    timed("B.pillar_canvas", pillars)

    def loss():
        losses = dbev.fgd.fgd_distill_loss(
            hp.teacher, hp.student, hp.boxes, bench.DISTILL_PARAMS, bench.TRAIN_CFG,
            channel_adaptation=hp.adapt, spatial_adaptation=hp.spatial, heatmaps=hp.d_gt_hm, teacher_heatmaps=hp.teacher_logit, epoch=1)
        state["total"] = losses["kd_fg_feat_loss"] + losses["kd_bg_feat_loss"] + losses["kd_spatial_loss"] \
            + losses["kd_fp_bg_feat_loss"]

    def lbwd():
        state["total"].backward(retain_graph=True)
        hp.student.grad = None
        hp.adapt.zero_grad(set_to_none=True)
        hp.spatial.zero_grad(set_to_none=True)
    timed("C.adapt+loss_fwd", loss)
    timed("C.loss_bwd", lbwd)
    return res


def main():
    dev = torch.device("cuda:0")
    hp = bench.HotPath(dev, 0)
    for _ in range(5):
        hp.step(False)
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n):
        hp.step(False)
    t1 = time.perf_counter()
    b.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = {"cpu_issue_ms_per_step": (t1 - t0) / n * 1e3, "gpu_ms_per_step": a.elapsed_time(b) / n,
           "wall_ms_per_step": (t2 - t0) / n * 1e3}
    out["phases"] = phases(hp)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            hp.step(False)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    busy = sum(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total for e in ev) / 3
    out["gpu_busy_us_per_step"] = busy
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        d = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        k = e.name.replace("(anonymous namespace)::", "").replace("void ", "")[:70]
        agg[k][0] += 1
        agg[k][1] += d
    out["top"] = [(k, c / 3, round(t / 3, 1)) for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]]
    print(json.dumps({k: v for k, v in out.items() if k not in ("top", "phases")}))
    for k, v in out["phases"].items():
        print("%-18s cpu issue %8.4f ms   gpu %8.4f ms" % (k, v["cpu_issue_ms"], v["gpu_ms"]))
    for k, c, t in out["top"]:
        print("%8.1f us  x%-4.1f %s" % (t, c, k))


if __name__ == "__main__":
    main()
