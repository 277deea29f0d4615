"""CPU oracle for the teacher pillar path (numpy).

TEST INFRASTRUCTURE ONLY — never imported by ``distill-bev_b200/``.

Restates (paths relative to the reference checkout):
  pillar_encode   DynamicPillarFeatureNet.forward, eval mode, one PFN layer
                  mmdet3d/models/voxel_encoders/pillar_encoder.py:282-338
                  (cluster mean :303-304, map_voxel_center_to_point :243-280,
                   f_cluster :307, f_center :312-317, Linear+BN1d+ReLU :221-233,
                   pfn_scatter max :330; DynamicScatter voxel order = lexicographic
                   (b, z, y, x), scatter_points.py:86-100 + unique_dim)
  pillar_scatter  PointPillarsScatter.forward_batch
                  mmdet3d/models/middle_encoders/pillar_scatter.py:62-102
Parity pin: tests/golden/pillar_small.npz is produced by the unmodified reference classes
(tools/make_golden.py, DynamicScatter's CUDA kernel replaced by its own host sequence on torch
CPU); tests/test_oracle_pillar.py checks this file against it.
"""
import numpy as np

from . import voxel_oracle

F32 = np.float32


def pillar_encode(points, coors, weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, voxel_size,
                  point_cloud_range):
    """points [N, F] f32, coors [N, 4] int32 (b, z, y, x; -1 = dropped) -> voxel_feats [M, nout],
    voxel_coors [M, 4]. float32 per-point math, like the reference."""
    pts = np.asarray(points, dtype=F32)
    co = np.asarray(coors, dtype=np.int32)
    mean, vcoors, cmap, cnt = voxel_oracle.dynamic_scatter(pts, co, "mean")
    valid = cmap >= 0
    pm = np.zeros_like(pts)
    pm[valid] = mean[cmap[valid]]
    # invalid points read canvas index of coordinate -1 in the reference (garbage that never reaches
    # the output because pfn_scatter drops them); keep them finite here
    f_cluster = (pts[:, :3] - pm[:, :3]).astype(F32)
    vx, vy = F32(voxel_size[0]), F32(voxel_size[1])
    x_off = F32(float(voxel_size[0]) / 2 + float(point_cloud_range[0]))
    y_off = F32(float(voxel_size[1]) / 2 + float(point_cloud_range[1]))
    f_center = np.stack([pts[:, 0] - ((co[:, 3].astype(F32) * vx).astype(F32) + x_off).astype(F32),
                         pts[:, 1] - ((co[:, 2].astype(F32) * vy).astype(F32) + y_off).astype(F32)], 1)
    feats = np.concatenate([pts, f_cluster, f_center.astype(F32)], axis=1).astype(np.float64)
    y = feats @ np.asarray(weight, dtype=np.float64).T
    scale = np.asarray(bn_weight, np.float64) / np.sqrt(np.asarray(bn_var, np.float64) + float(bn_eps))
    y = (y - np.asarray(bn_mean, np.float64)) * scale + np.asarray(bn_bias, np.float64)
    y = np.maximum(y, 0.0).astype(F32)
    vf, vc, _, _ = voxel_oracle.dynamic_scatter(y, co, "max")
    assert np.array_equal(vc, vcoors)
    return vf, vc


def pillar_scatter(voxel_feats, coors, batch_size, ny, nx):
    """-> canvas [B, C, ny, nx] float32."""
    vf = np.asarray(voxel_feats, dtype=F32)
    co = np.asarray(coors).astype(np.int64)
    canvas = np.zeros((batch_size, vf.shape[1], ny, nx), dtype=F32)
    canvas[co[:, 0], :, co[:, 2], co[:, 3]] = vf
    return canvas


def hard_pillar_encode(voxels, num_points, coors, weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, voxel_size,
                       point_cloud_range, legacy=False):
    """PillarFeatureNet.forward in eval mode (pillar_encoder.py:95-162; PFNLayer utils.py:107-181, one layer, max):
    voxels [M, T, F] f32 (padded rows as hard_voxelize leaves them), num_points [M], coors [M, 4] (b, z, y, x)
    -> [M, nout]. The mean divides the sum over ALL T rows by num_points (:121-124); padded rows are zeroed before
    the linear layer (:149-153) and therefore enter the maximum as relu(BN(0)); legacy=True: f_center is a view of the
    raw features, so x / y of the raw block hold the centre offsets too (:132-139)."""
    v = np.asarray(voxels, dtype=F32).copy()
    n = np.asarray(num_points).astype(np.int64)
    co = np.asarray(coors).astype(np.int64)
    M, T, F = v.shape
    mean = (v[:, :, :3].sum(axis=1, keepdims=True, dtype=F32) / n.astype(F32).reshape(-1, 1, 1)).astype(F32)
    f_cluster = (v[:, :, :3] - mean).astype(F32)
    vx, vy = F32(voxel_size[0]), F32(voxel_size[1])
    x_off = F32(float(voxel_size[0]) / 2 + float(point_cloud_range[0]))
    y_off = F32(float(voxel_size[1]) / 2 + float(point_cloud_range[1]))
    cxs = ((co[:, 3].astype(F32) * vx).astype(F32) + x_off).astype(F32).reshape(-1, 1)
    cys = ((co[:, 2].astype(F32) * vy).astype(F32) + y_off).astype(F32).reshape(-1, 1)
    f_center = np.stack([v[:, :, 0] - cxs, v[:, :, 1] - cys], 2).astype(F32)
    raw = v.copy()
    if legacy:
        raw[:, :, :2] = f_center
    feats = np.concatenate([raw, f_cluster, f_center], axis=2)
    mask = (np.arange(T).reshape(1, T) < n.reshape(-1, 1)).astype(F32)[:, :, None]
    feats = (feats * mask).astype(np.float64)
    y = feats @ np.asarray(weight, dtype=np.float64).T
    scale = np.asarray(bn_weight, np.float64) / np.sqrt(np.asarray(bn_var, np.float64) + float(bn_eps))
    y = (y - np.asarray(bn_mean, np.float64)) * scale + np.asarray(bn_bias, np.float64)
    return np.maximum(y, 0.0).max(axis=1).astype(F32)
