"""Run-to-run difference of the training step's gradients (same weights, same inputs), per tensor, for the four
combinations of {side stream on/off} x {sort-free / sorted splat}."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for side in ("1", "0"):
        for sorted_splat in (False, True):
            os.environ["DBEV_BENCH_SIDE_STREAM"] = side
            hp = bench.HotPath(dev, seed=1000)
            hp.sorted_splat = sorted_splat
            packed = hp.dbev.fgd.PackedBoxes(hp.d_boxes, hp.d_box_offs, hp.max_boxes)
            names = []
            for mod_name, m in (("enc", hp.student_net), ("adapt", hp.adapt), ("spatial", hp.spatial)):
                names += [mod_name + "." + k for k, _ in m.named_parameters()]
            runs = []
            for _ in range(3):
                for p in hp.trainable:
                    p.grad = None
                hp._forward_backward(hp.d_calib, hp.d_points, hp.d_labels, packed)
                hp.dbev.conv_train.join_side_stream(dev)
                torch.cuda.synchronize()
                runs.append([p.grad.detach().clone() for p in hp.trainable])
            worst = []
            for n, a, b, c in zip(names, *runs):
                s = float(a.abs().max().clamp_min(1e-30))
                worst.append((max(float((a - b).abs().max()), float((a - c).abs().max())) / s, n, s))
            worst.sort(reverse=True)
            print("side", side, "sorted_splat", sorted_splat, "top:", [(round(w, 6), n, "%.2e" % s) for w, n, s in worst[:5]])
            del hp
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
