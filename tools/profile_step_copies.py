"""Which torch ops inside the bench step launch copy / elementwise kernels (shapes + call sites)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    hp = bench.HotPath(dev, seed=0)
    for _ in range(3):
        hp.step(False)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        hp.step(False)
        torch.cuda.synchronize()
    rows = []
    for e in prof.events():
        if e.device_time_total > 0 and e.name.startswith("aten::"):
            stack = [s for s in e.stack if "distill" in s or "bench.py" in s][:3]
            rows.append((e.device_time_total, e.name, str(e.input_shapes)[:90], " <- ".join(x.split("/")[-1][:60] for x in stack)))
    rows.sort(reverse=True)
    for r in rows[:40]:
        print("%8.1f us  %-28s %s  %s" % r)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
