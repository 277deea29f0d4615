"""Timeline of ONE replay of the bench step's CUDA graph from CUPTI kernel records (torch.profiler): per stream the busy
time and the idle gaps, for every kernel its start offset, duration and how many other kernels ran beside it. Shows
what the critical path of the step is made of (the ncu launch list gives serialised, cold-cache durations only).
Writes gpurun_out/step_timeline.txt."""
import collections
import json
import os
import re
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return name[:64]


def main():
    dev = torch.device("cuda:0")
    hp = bench.HotPath(dev, seed=1000)
    hp.enable_graph()
    for _ in range(5):
        hp.step(False)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            hp.step(False)
            torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "dbev_step_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    # split into replays at the largest gaps between consecutive kernels
    gaps = sorted(((ev[i + 1]["ts"] - (ev[i]["ts"] + ev[i]["dur"]), i) for i in range(len(ev) - 1)), reverse=True)[:2]
    cuts = sorted(i for _, i in gaps)
    step = ev[cuts[0] + 1:cuts[1] + 1]           # the middle replay
    t0 = min(e["ts"] for e in step)
    t1 = max(e["ts"] + e["dur"] for e in step)
    out = []
    out.append("replay: %.1f us wall, %d kernels" % (t1 - t0, len(step)))
    by_stream = collections.OrderedDict()
    for e in step:
        by_stream.setdefault(e["args"].get("stream"), []).append(e)
    main_stream = max(by_stream, key=lambda s: len(by_stream[s]))
    for s, es in by_stream.items():
        busy = sum(e["dur"] for e in es)
        first, last = es[0]["ts"] - t0, es[-1]["ts"] + es[-1]["dur"] - t0
        out.append("stream %s%s: %3d kernels, busy %.1f us, active window %.1f .. %.1f us" %
                   (s, " (main)" if s == main_stream else "", len(es), busy, first, last))
    # main-stream idle gaps
    es = by_stream[main_stream]
    idle = [(es[i + 1]["ts"] - (es[i]["ts"] + es[i]["dur"]), short(es[i]["name"]), short(es[i + 1]["name"])) for i in range(len(es) - 1)]
    out.append("main stream: sum of kernel durations %.1f us, sum of idle gaps %.1f us (%d gaps > 3 us)" %
               (sum(e["dur"] for e in es), sum(g for g, _, _ in idle), sum(1 for g, _, _ in idle if g > 3)))
    for g, a, b in sorted(idle, reverse=True)[:25]:
        out.append("   gap %7.1f us  after %-50s before %s" % (g, a[:50], b))
    # time with k streams active
    edges = []
    for e in step:
        edges.append((e["ts"], 1)), edges.append((e["ts"] + e["dur"], -1))
    edges.sort()
    act, last_t, hist = 0, t0, collections.Counter()
    for t, d in edges:
        hist[act] += t - last_t
        act, last_t = act + d, t
    out.append("time by number of kernels in flight: " + ", ".join("%d: %.0f us" % (k, v) for k, v in sorted(hist.items())))
    # per-kernel-name totals inside the replay (durations under contention)
    agg = collections.OrderedDict()
    for e in step:
        a = agg.setdefault((short(e["name"]), e["args"].get("stream") == main_stream), [0, 0.0])
        a[0] += 1
        a[1] += e["dur"]
    out.append("kernel totals in the replay (under contention):")
    for (k, is_main), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        out.append("   %-66s %s %3d %8.1f us" % (k, "main" if is_main else "side", n, t))
    if "-v" in sys.argv:
        out.append("sequence (start us, dur us, stream, kernel):")
        for e in step:
            out.append("   %8.1f %7.1f %3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), short(e["name"])))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "step_timeline.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[:90]))


if __name__ == "__main__":
    main()
