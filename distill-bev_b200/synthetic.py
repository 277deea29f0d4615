"""Seeded synthetic nuScenes-shaped inputs for tests and bench.

Pure input generation on the CPU (numpy); no part of the hot path. There is
no network / dataset in the build or GPU containers, so every measurement
uses these tensors and says ``"data": "synthetic"``.
"""
import math

import numpy as np

NUSC_GRID = dict(xbound=[-51.2, 51.2, 0.8], ybound=[-51.2, 51.2, 0.8],
                 zbound=[-10.0, 10.0, 20.0], dbound=[1.0, 60.0, 1.0])
NUSC_INPUT_SIZE = (256, 704)
NUSC_SRC_SIZE = (900, 1600)


def grid_config(bev=128, dstep=1.0):
    """xbound/ybound for a bev x bev grid over +-51.2 m; dbound [1, 60, dstep]."""
    step = 102.4 / bev
    return dict(xbound=[-51.2, 51.2, step], ybound=[-51.2, 51.2, step],
                zbound=[-10.0, 10.0, 20.0], dbound=[1.0, 60.0, dstep])


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def make_calibration(batch, n_cams=6, seed=0, input_size=NUSC_INPUT_SIZE, src_size=NUSC_SRC_SIZE,
                     augment=True):
    """rots, trans, intrins, post_rots, post_trans as float32 arrays [B, N, ...].

    Cameras look outward every 360/n_cams degrees; intrinsics are nuScenes-like
    (f = 1266 px, principal point (816, 491) at 1600x900); the image-space
    augmentation (resize / crop / flip / rotate) follows the BEVDet recipe the
    reference loader applies (mmdet3d/datasets/pipelines/loading.py:171-227).
    """
    rng = np.random.RandomState(seed)
    fH, fW = input_size
    H, W = src_size
    cam2ego0 = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    rots = np.zeros((batch, n_cams, 3, 3))
    trans = np.zeros((batch, n_cams, 3))
    intrins = np.zeros((batch, n_cams, 3, 3))
    post_rots = np.zeros((batch, n_cams, 3, 3))
    post_trans = np.zeros((batch, n_cams, 3))
    for b in range(batch):
        for k in range(n_cams):
            yaw = 2.0 * math.pi * k / n_cams + (rng.uniform(-0.02, 0.02) if augment else 0.0)
            rots[b, k] = _rot_z(yaw) @ cam2ego0
            trans[b, k] = rng.uniform(-1, 1, 3) * np.array([1.5, 1.5, 0.5]) + np.array([0, 0, 1.5])
            intrins[b, k] = np.array([[1266.0, 0, 816.0], [0, 1266.0, 491.0], [0, 0, 1.0]])
            resize = float(fW) / float(W) + (rng.uniform(-0.06, 0.11) if augment else 0.0)
            newW, newH = int(W * resize), int(H * resize)
            crop_h = newH - fH
            crop_w = int(rng.uniform(0, 1) * max(0, newW - fW)) if augment else max(0, newW - fW) // 2
            flip = bool(rng.randint(0, 2)) if augment else False
            rotate = rng.uniform(-5.4, 5.4) if augment else 0.0
            A = np.eye(2) * resize
            t = -np.array([crop_w, crop_h], dtype=np.float64)
            if flip:
                F = np.array([[-1.0, 0.0], [0.0, 1.0]])
                A = F @ A
                t = F @ t + np.array([fW, 0.0])
            th = rotate / 180.0 * math.pi
            R = np.array([[math.cos(th), math.sin(th)], [-math.sin(th), math.cos(th)]])
            c = np.array([fW, fH]) / 2.0
            A = R @ A
            t = R @ (t - c) + c
            post_rots[b, k] = np.eye(3)
            post_rots[b, k, :2, :2] = A
            post_trans[b, k, :2] = t
    f32 = np.float32
    return (rots.astype(f32), trans.astype(f32), intrins.astype(f32), post_rots.astype(f32),
            post_trans.astype(f32))


def make_frustum_feats(n, channels, seed=0):
    """Lifted frustum features (softmax-depth x image feature): U(0, 1) float32."""
    rng = np.random.RandomState(seed)
    return rng.random_sample((n, channels)).astype(np.float32)


def make_lidar(batch, n_points=30000, seed=0, num_features=5):
    """List of [Np, F] float32 clouds strictly inside the nuScenes range.

    x, y: 70 % N(0, 15^2) + 30 % U(-51.2, 51.2); z ~ N(-1, 1) clipped to (-5, 3);
    intensity U(0, 255); time lag in {0, 0.05, ..., 0.45}; extra features U(0, 1).
    """
    clouds = []
    for b in range(batch):
        rng = np.random.RandomState(seed * 1000 + b)
        near = rng.random_sample(n_points) < 0.7
        xy = np.where(near[:, None], rng.normal(0, 15.0, (n_points, 2)),
                      rng.uniform(-51.2, 51.2, (n_points, 2)))
        xy = np.clip(xy, -51.19, 51.19)
        z = np.clip(rng.normal(-1.0, 1.0, n_points), -4.99, 2.99)
        pts = np.zeros((n_points, num_features), dtype=np.float32)
        pts[:, 0:2] = xy
        pts[:, 2] = z
        if num_features > 3:
            pts[:, 3] = rng.uniform(0, 255, n_points)
        if num_features > 4:
            pts[:, 4] = rng.randint(0, 10, n_points) * 0.05
        if num_features > 5:
            pts[:, 5:] = rng.random_sample((n_points, num_features - 5))
        clouds.append(pts)
    return clouds


_CLASS_SIZES = np.array([  # (w, l, h) priors for the 10 nuScenes classes
    [1.95, 4.60, 1.73], [2.50, 6.90, 2.84], [2.80, 6.40, 3.20], [2.95, 11.0, 3.47],
    [2.90, 12.3, 3.80], [2.50, 0.50, 0.98], [0.77, 2.10, 1.47], [0.60, 1.70, 1.28],
    [0.67, 0.73, 1.77], [0.41, 0.41, 1.07]])


def make_gt_boxes(batch, seed=0, min_boxes=5, max_boxes=60):
    """Per sample: boxes [M, 9] (x, y, z_bottom, w, l, h, yaw, vx, vy) float32 and labels [M]."""
    out = []
    for b in range(batch):
        rng = np.random.RandomState(seed * 977 + b)
        m = rng.randint(min_boxes, max_boxes + 1)
        labels = rng.randint(0, 10, m)
        size = _CLASS_SIZES[labels] * rng.uniform(0.85, 1.15, (m, 3))
        xy = rng.uniform(-48.0, 48.0, (m, 2))
        z = rng.uniform(-2.5, -0.5, (m, 1))
        yaw = rng.uniform(-math.pi, math.pi, (m, 1))
        vel = rng.normal(0, 2.0, (m, 2))
        boxes = np.concatenate([xy, z, size, yaw, vel], axis=1).astype(np.float32)
        out.append((boxes, labels.astype(np.int64)))
    return out


def make_lidar_scene(batch, n_points=240000, seed=0, num_features=5):
    """LiDAR-like clouds for the SPARSE teacher benchmarks: points lie on surfaces (ground rings of
    a 32-beam sensor accumulated over sweeps, plus vertical object / wall faces), so that occupied
    voxels have occupied neighbours the way real sweeps do — make_lidar()'s volumetric Gaussian
    gives a submanifold rulebook with almost no pairs at 6.4 cm voxels."""
    clouds = []
    for b in range(batch):
        rng = np.random.RandomState(seed * 1000 + 17 + b)
        n_ground = int(n_points * 0.6)
        beams = np.deg2rad(rng.choice(np.linspace(-30.0, -1.5, 24), n_ground))
        r = np.clip(1.84 / np.tan(-beams) + rng.normal(0, 0.03, n_ground), 1.0, 72.0)
        az = rng.uniform(0, 2 * np.pi, n_ground)
        ground = np.stack([r * np.cos(az), r * np.sin(az), -1.84 + rng.normal(0, 0.02, n_ground)], 1)
        n_obj = n_points - n_ground
        n_faces = 160
        centre = rng.uniform(-45, 45, (n_faces, 2))
        yaw = rng.uniform(0, np.pi, n_faces)
        length = rng.uniform(1.5, 12.0, n_faces)
        height = rng.uniform(1.0, 3.5, n_faces)
        f = rng.randint(0, n_faces, n_obj)
        u = rng.uniform(-0.5, 0.5, n_obj) * length[f]
        obj = np.stack([centre[f, 0] + u * np.cos(yaw[f]), centre[f, 1] + u * np.sin(yaw[f]),
                        -1.84 + rng.uniform(0, 1, n_obj) * height[f]], 1)
        obj += rng.normal(0, 0.015, obj.shape)
        xyz = np.concatenate([ground, obj], 0)
        xyz[:, :2] = np.clip(xyz[:, :2], -51.19, 51.19)
        xyz[:, 2] = np.clip(xyz[:, 2], -4.99, 2.99)
        xyz = xyz[rng.permutation(n_points)]
        pts = np.zeros((n_points, num_features), dtype=np.float32)
        pts[:, :3] = xyz
        if num_features > 3:
            pts[:, 3] = rng.uniform(0, 255, n_points)
        if num_features > 4:
            pts[:, 4] = rng.randint(0, 10, n_points) * 0.05
        if num_features > 5:
            pts[:, 5:] = rng.random_sample((n_points, num_features - 5))
        clouds.append(pts)
    return clouds
