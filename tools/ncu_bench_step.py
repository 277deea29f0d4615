"""One eager bench step under cudaProfilerStart/Stop (for `ncu --profile-from-start off`): the launch list of the
training step (profiles/r02_launches_bench_step.json is its summary)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    hp = bench.HotPath(dev, seed=bench.rank_seed(0))
    for _ in range(3):
        hp.step(False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    hp.step(False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
