"""CenterHead training targets on the device: mirror of
``CenterHead.get_targets`` (mmdet3d/models/dense_heads/centerpoint_head.py:400-445, 447-611).

The reference builds the targets of every sample, task and object in Python (radius, centre, a
slice-max of the Gaussian and a torch.cat of ten 0-d tensors per object); here ONE kernel per batch.
"""
import torch

from .. import _lib
from .distill import fgd as _fgd


class CenterHeadTargets(object):
    """``tasks`` = [dict(num_class=..., class_names=[...]), ...] as in the CenterHead config;
    ``train_cfg`` = pts_bbox_head.train_cfg (grid_size, point_cloud_range, voxel_size, out_size_factor,
    max_objs, dense_reg, gaussian_overlap, min_radius)."""

    def __init__(self, tasks, train_cfg, norm_bbox=True):
        self.class_names = [list(t["class_names"]) for t in tasks]
        self.train_cfg = train_cfg
        self.norm_bbox = norm_bbox
        self.class_task, self.class_in_task = [], []
        for ti, names in enumerate(self.class_names):
            for ci in range(len(names)):
                self.class_task.append(ti)
                self.class_in_task.append(ci)

    def get_targets(self, gt_bboxes_3d, gt_labels_3d, device=None):
        """-> (heatmaps, anno_boxes, inds, masks): lists over tasks of [B, K_t, H, W] float32,
        [B, max_objs, 10] float32, [B, max_objs] int64, [B, max_objs] uint8 — what the reference
        returns after its transposes / stacks (:435-445). Also returns nothing else; the full
        class heat map [B, num_classes, H, W] is ``self.last_heatmap``."""
        lib = _lib.load()
        cfg = self.train_cfg
        if device is None:
            device = gt_labels_3d.device if isinstance(gt_labels_3d, torch.Tensor) else gt_labels_3d[0].device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("get_targets: device must be CUDA (no CPU path)")
        boxes, offs, _ = _fgd.pack_boxes(gt_bboxes_3d, device)
        if isinstance(gt_labels_3d, torch.Tensor):
            # already packed on the device (same order as the packed boxes): CUDA-graph friendly
            labels = _lib.require_cuda(gt_labels_3d, "gt_labels_3d", torch.int32).contiguous()
            B = len(gt_bboxes_3d)
        else:
            labels = torch.cat([torch.as_tensor(l).reshape(-1) for l in gt_labels_3d]).to(torch.int32)
            if labels.numel() == 0:
                labels = torch.zeros((1,), dtype=torch.int32)
            labels = labels.contiguous().to(device) if labels.is_cuda else _lib.h2d_async(labels, device)
            B = len(gt_labels_3d)
        nc, nt = len(self.class_task), len(self.class_names)
        max_objs = int(cfg["max_objs"] * cfg["dense_reg"])
        osf = cfg["out_size_factor"]
        W, H = int(cfg["grid_size"][0]) // osf, int(cfg["grid_size"][1]) // osf
        heat = torch.empty((B, nc, H, W), dtype=torch.float32, device=device)
        anno = torch.empty((B, nt, max_objs, 10), dtype=torch.float32, device=device)
        ind = torch.empty((B, nt, max_objs), dtype=torch.int64, device=device)
        mask = torch.empty((B, nt, max_objs), dtype=torch.uint8, device=device)
        vs, pcr = cfg["voxel_size"], cfg["point_cloud_range"]
        f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
        with torch.cuda.device(device):
            rc = lib.dbev_center_targets(
                _lib.ptr(boxes), boxes.shape[1], _lib.ptr(labels), _lib.ptr(offs), B,
                _lib.host_ints(self.class_task), _lib.host_ints(self.class_in_task), nc, nt, max_objs, H, W,
                f32(vs[0]), f32(vs[1]), float(osf), f32(pcr[0]), f32(pcr[1]), float(cfg["gaussian_overlap"]),
                int(cfg["min_radius"]), int(bool(self.norm_bbox)), _lib.ptr(heat), _lib.ptr(anno), _lib.ptr(ind),
                _lib.ptr(mask), _lib.stream_ptr(device))
        _lib.check(rc, "dbev_center_targets")
        self.last_heatmap = heat
        heatmaps, c0 = [], 0
        for names in self.class_names:
            heatmaps.append(heat[:, c0:c0 + len(names)])
            c0 += len(names)
        return (heatmaps, [anno[:, t] for t in range(nt)], [ind[:, t] for t in range(nt)],
                [mask[:, t] for t in range(nt)])
