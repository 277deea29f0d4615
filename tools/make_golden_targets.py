"""Golden vectors for the CenterHead targets, produced by EXECUTING the unmodified
CenterHead.get_targets / get_targets_single (mmdet3d/models/dense_heads/centerpoint_head.py:400-611)
with the reference's own gaussian helpers (mmdet3d/core/utils/gaussian.py); mmdet's multi_apply
(third party) is restated.    python tools/make_golden_targets.py -> tests/golden/center_targets.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class _Boxes(object):
    def __init__(self, t):
        self.tensor = t

    @property
    def gravity_center(self):          # LiDARInstance3DBoxes.gravity_center (lidar_box3d.py)
        gc = self.tensor[:, :3].clone()
        gc[:, 2] = self.tensor[:, 2] + self.tensor[:, 5] * 0.5
        return gc


def multi_apply(func, *args, **kwargs):
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def main():
    gauss = ref_import.load_ref_module("ref_gaussian", "mmdet3d/core/utils/gaussian.py")
    methods, _ = ref_import.load_fgd_methods(
        names=("get_targets", "get_targets_single"), cls_name="CenterHead",
        relpath="mmdet3d/models/dense_heads/centerpoint_head.py",
        extra_ns=dict(draw_heatmap_gaussian=gauss.draw_heatmap_gaussian, gaussian_radius=gauss.gaussian_radius,
                      multi_apply=multi_apply))
    tasks = [dict(num_class=1, class_names=["car"]), dict(num_class=2, class_names=["truck", "construction_vehicle"]),
             dict(num_class=2, class_names=["bus", "trailer"]), dict(num_class=1, class_names=["barrier"]),
             dict(num_class=2, class_names=["motorcycle", "bicycle"]),
             dict(num_class=2, class_names=["pedestrian", "traffic_cone"])]
    train_cfg = dict(grid_size=[256, 256, 40], point_cloud_range=[-25.6, -25.6, -5.0, 25.6, 25.6, 3.0],
                     voxel_size=[0.2, 0.2, 8.0], out_size_factor=4, dense_reg=1, gaussian_overlap=0.1, max_objs=40,
                     min_radius=2)

    class Fake(object):
        get_targets = methods["get_targets"]
        get_targets_single = methods["get_targets_single"]
    me = Fake()
    me.class_names = [t["class_names"] for t in tasks]
    me.train_cfg = train_cfg
    me.task_heads = [None] * len(tasks)
    me.norm_bbox = True
    rs = np.random.RandomState(5)
    boxes, labels = [], []
    for m in (70, 0, 23):
        b = np.zeros((m, 9), np.float32)
        b[:, :2] = rs.uniform(-27, 27, (m, 2))          # some centres fall outside the range
        b[:, 2] = rs.uniform(-3, 0, m)
        b[:, 3:6] = rs.uniform(0.4, 9.0, (m, 3))
        b[:, 6] = rs.uniform(-3.14, 3.14, m)
        b[:, 7:] = rs.uniform(-5, 5, (m, 2))
        if m:
            b[0, 3] = 0.0                               # zero width: skipped (:526)
        l = rs.randint(0, 10, m)
        l[: m // 2] = 0                                  # many cars: exceeds max_objs=40 in task 0
        boxes.append(b)
        labels.append(l.astype(np.int64))
    hm, ab, ind, mk = me.get_targets([_Boxes(torch.from_numpy(b)) for b in boxes], [torch.from_numpy(l) for l in labels])
    out = dict(n=np.array([len(b) for b in boxes]), boxes=np.concatenate(boxes), labels=np.concatenate(labels))
    for t in range(len(tasks)):
        out["hm%d" % t], out["anno%d" % t] = hm[t].numpy(), ab[t].numpy()
        out["ind%d" % t], out["mask%d" % t] = ind[t].numpy(), mk[t].numpy()
    print("targets: heatmap sums", [float(h.sum()) for h in hm], "valid", [int(m.sum()) for m in mk])
    path = os.path.join(GOLDEN, "center_targets.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
