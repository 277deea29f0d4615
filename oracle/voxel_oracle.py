"""numpy front-end of the C voxelization oracle (oracle/c/voxel_oracle.c) and loader of
the reference's own CPU extension (oracle/_ref/, built by oracle/build_oracle.py).

TEST INFRASTRUCTURE ONLY — see the header of voxel_oracle.c for the reference
file:line each function follows and for how parity is pinned.
"""
import ctypes
import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            import sys
            subprocess.run([sys.executable, os.path.join(_HERE, "build_oracle.py")], check=True)
        _LIB = ctypes.CDLL(path)
        _LIB.vo_hard_voxelize.restype = ctypes.c_int
        _LIB.vo_dynamic_scatter.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def grid_size(voxel_size, coors_range):
    g = np.zeros(3, dtype=np.int32)
    _lib().vo_grid_size(_p(_f32(voxel_size)), _p(_f32(coors_range)), _p(g))
    return tuple(int(v) for v in g)


def dynamic_voxelize(points, voxel_size, coors_range):
    pts = _f32(points)
    n, f = pts.shape
    coors = np.zeros((n, 3), dtype=np.int32)
    _lib().vo_dynamic_voxelize(_p(pts), n, f, _p(_f32(voxel_size)), _p(_f32(coors_range)), _p(coors))
    return coors


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels):
    """-> (voxels[M, max_points, F], coors[M, 3], num_points_per_voxel[M])."""
    pts = _f32(points)
    n, f = pts.shape
    voxels = np.zeros((max_voxels, max_points, f), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    m = _lib().vo_hard_voxelize(_p(pts), n, f, _p(_f32(voxel_size)), _p(_f32(coors_range)),
                                int(max_points), int(max_voxels), _p(voxels), _p(coors), _p(num))
    return voxels[:m], coors[:m], num[:m]


_RED = {"sum": 0, "mean": 1, "max": 2}


def dynamic_scatter(feats, coors, reduce_type):
    """-> (reduced[M, C], out_coors[M, ncol], coors_map[N], reduce_count[M])."""
    ft = _f32(feats)
    co = np.ascontiguousarray(coors, dtype=np.int32)
    n, c = ft.shape
    ncol = co.shape[1]
    reduced = np.zeros((max(n, 1), c), dtype=np.float32)
    out_coors = np.zeros((max(n, 1), ncol), dtype=np.int32)
    cmap = np.zeros((max(n, 1),), dtype=np.int32)
    cnt = np.zeros((max(n, 1),), dtype=np.int32)
    m = _lib().vo_dynamic_scatter(_p(ft), _p(co), n, c, ncol, _RED[reduce_type], _p(reduced),
                                  _p(out_coors), _p(cmap), _p(cnt))
    return reduced[:m], out_coors[:m], cmap[:n], cnt[:m]


def dynamic_scatter_backward(grad_reduced, feats, reduced, coors_map, reduce_count, reduce_type):
    ft = _f32(feats)
    n, c = ft.shape
    m = reduced.shape[0]
    g = np.zeros((n, c), dtype=np.float32)
    _lib().vo_dynamic_scatter_backward(
        _p(_f32(grad_reduced)), _p(ft), _p(_f32(reduced)),
        _p(np.ascontiguousarray(coors_map, dtype=np.int32)),
        _p(np.ascontiguousarray(reduce_count, dtype=np.int32)), n, m, c, _RED[reduce_type], _p(g))
    return g


def load_reference_voxel_layer():
    """The reference's own CPU extension (hard_voxelize / dynamic_voxelize), or None."""
    cands = glob.glob(os.path.join(_HERE, "_ref", "ref_voxel_layer*.so"))
    if not cands:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("ref_voxel_layer", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
