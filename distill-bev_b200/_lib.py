"""ctypes binding of libdistill_bev_b200.so (the C-ABI in include/distill_bev_b200.h).

There is no CPU or PyTorch fallback anywhere in this package: if the shared
library is missing or a tensor is not on a CUDA device the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdistill_bev_b200.so")

_c_int = ctypes.c_int
_c_ll = ctypes.c_longlong
_c_size = ctypes.c_size_t
_ptr = ctypes.c_void_p
_fptr = ctypes.POINTER(ctypes.c_float)
_iptr = ctypes.POINTER(ctypes.c_int)

class FgdConfig(ctypes.Structure):
    """Mirror of struct dbev_fgd_config (include/distill_bev_b200.h)."""
    _fields_ = [("B", ctypes.c_int), ("C", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("spatial_t", ctypes.c_float), ("channel_t", ctypes.c_float),
                ("spatial_student_ratio", ctypes.c_float),
                ("w_fg", ctypes.c_float), ("w_bg", ctypes.c_float), ("w_fp", ctypes.c_float),
                ("w_channel", ctypes.c_float), ("w_spatial", ctypes.c_float),
                ("spatial_att", ctypes.c_int), ("spatial_mask", ctypes.c_int),
                ("channel_mask", ctypes.c_int), ("scale_mask", ctypes.c_int),
                ("use_fp", ctypes.c_int)]


_cfgp = ctypes.POINTER(FgdConfig)
_c_float = ctypes.c_float

# name -> (restype, argtypes); must list every symbol include/distill_bev_b200.h declares
SIGNATURES = {
    "dbev_abi_version": (_c_int, []),
    "dbev_last_error": (ctypes.c_char_p, []),
    "dbev_build_arch": (ctypes.c_char_p, []),
    "dbev_bev_pool_forward": (_c_int, [_c_int] * 7 + [_ptr] * 5 + [_c_int, _ptr]),
    "dbev_bev_pool_backward": (_c_int, [_c_int] * 7 + [_ptr] * 5 + [_c_int, _ptr]),
    "dbev_bev_plan_workspace_bytes": (_c_size, [_c_ll, _c_ll]),
    "dbev_bev_plan_max_items": (_c_ll, [_c_ll, _c_ll, _c_int, _c_int]),
    "dbev_bev_plan_from_geom": (_c_int, [_ptr, _c_ll, _c_int, _fptr, _fptr, _fptr, _iptr, _c_int,
                                         _c_int, _ptr, _ptr, _ptr, _ptr, _c_ll, _ptr, _ptr, _ptr,
                                         _c_size, _ptr]),
    "dbev_bev_pool_point_backward": (_c_int, [_ptr, _ptr, _c_ll, _c_int, _ptr, _ptr]),
    "dbev_lift_splat_forward": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr,
                                         _ptr, _c_int, _c_int, _c_int, _c_int, _c_ll, _c_ll, _c_ll,
                                         _ptr, _ptr]),
    "dbev_lift_splat_backward": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_ll, _c_int, _c_int, _c_int,
                                          _ptr, _ptr, _ptr]),
    "dbev_bev_point_cells": (_c_int, [_ptr, _c_ll, _c_int, _fptr, _fptr, _fptr, _iptr, _c_int, _ptr, _ptr]),
    "dbev_bev_point_cells_frames": (_c_int, [_ptr, _c_ll, _c_int, _c_int, _fptr, _fptr, _fptr, _iptr, _c_int, _ptr, _ptr]),
    "dbev_lift_splat_atomic_forward": (_c_int, [_ptr, _ptr, _ptr, _c_ll, _c_int, _c_int, _c_int, _c_ll, _ptr,
                                                _ptr]),
    "dbev_transpose_batched": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _ptr]),
    "dbev_bev_plan_from_coords": (_c_int, [_ptr, _c_int, _c_ll, _c_int, _c_int, _c_int, _c_int,
                                           _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_ll, _ptr,
                                           _ptr, _c_size, _ptr]),
    "dbev_bev_pool_gather_forward": (_c_int, [_ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int,
                                              _c_int, _c_int, _c_int, _c_ll, _c_ll, _c_ll, _ptr,
                                              _ptr]),
    "dbev_bev_pool_gather_backward": (_c_int, [_ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int,
                                               _c_int, _c_int, _c_int, _c_ll, _c_ll, _c_ll, _ptr,
                                               _ptr]),
    "dbev_voxel_grid_size": (_c_int, [_fptr, _fptr, _iptr]),
    "dbev_dynamic_voxelize": (_c_int, [_ptr, _c_int, _c_int, _fptr, _fptr, _ptr, _ptr]),
    "dbev_hard_voxelize_workspace_bytes": (_c_size, [_c_ll]),
    "dbev_hard_voxelize": (_c_int, [_ptr, _c_int, _c_int, _fptr, _fptr, _c_int, _c_int, _ptr, _ptr,
                                    _ptr, _ptr, _ptr, _c_size, _ptr]),
    "dbev_dynamic_scatter_workspace_bytes": (_c_size, [_c_ll]),
    "dbev_dynamic_scatter_forward": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _iptr, _c_int,
                                              _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_size, _ptr]),
    "dbev_dynamic_scatter_backward": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c_ll, _c_ll, _c_int,
                                               _c_int, _ptr, _ptr, _ptr]),
    "dbev_fgd_foreground_mask": (_c_int, [_ptr, _c_int, _ptr, _c_int, _c_int, _c_int, _c_int, _c_float,
                                          _c_float, _c_float, _c_float, _c_float, _c_int, _c_int,
                                          _ptr, _ptr, _ptr, _ptr]),
    "dbev_heatmap_class_max": (_c_int, [_ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr]),
    "dbev_fgd_fp_dfs_workspace_bytes": (_c_size, [_c_int, _c_int, _c_int]),
    "dbev_fgd_fp_dfs_scale": (_c_int, [_ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _c_size, _ptr]),
    "dbev_fgd_fp_mask": (_c_int, [_ptr, _c_int, _ptr, _c_int, _ptr, _c_int, _ptr, _c_int, _c_int,
                                  _c_int, _c_float, _c_float, _ptr, _ptr, _ptr]),
    "dbev_fgd_state_bytes": (_c_size, [_cfgp]),
    "dbev_fgd_loss_forward": (_c_int, [_cfgp, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                       _ptr, _c_size, _ptr, _ptr]),
    "dbev_fgd_loss_backward": (_c_int, [_cfgp, _ptr, _ptr, _ptr, _ptr, _ptr, _c_size, _ptr, _ptr,
                                        _ptr, _ptr, _ptr, _ptr]),
    "dbev_fgd_adapt_supported": (_c_int, [_cfgp, _c_int]),
    "dbev_fgd_adapt_loss_forward": (_c_int, [_cfgp, _ptr, _c_int] + [_ptr] * 11 + [_c_size, _ptr, _ptr]),
    "dbev_fgd_adapt_loss_backward": (_c_int, [_cfgp, _ptr, _c_int] + [_ptr] * 6 + [_c_size] + [_ptr] * 6),
    "dbev_pillar_encode_workspace_bytes": (_c_size, [_c_ll]),
    "dbev_pillar_encode": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _fptr, _fptr, _c_float,
                                    _c_float, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                    _c_size, _ptr]),
    "dbev_pillar_canvas": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _fptr, _fptr, _c_float,
                                    _c_float, _ptr, _c_int, _ptr, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr,
                                    _c_size, _ptr]),
    "dbev_pillar_scatter": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                     _c_int, _ptr, _ptr]),
    "dbev_lss_geometry": (_c_int, [_ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr,
                                   _ptr]),
    "dbev_adapt_conv1x1_forward": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr]),
    "dbev_spconv_max_out": (_c_ll, [_c_ll, _iptr]),
    "dbev_spconv_workspace_bytes": (_c_size, [_c_ll, _c_ll]),
    "dbev_spconv_table": (_c_int, [_ptr, _c_int, _ptr, _c_int, _iptr, _ptr, _ptr, _c_size, _ptr]),
    "dbev_spconv_out_candidates": (_c_int, [_ptr, _c_int, _iptr, _ptr, _c_ll, _ptr, _ptr, _c_size,
                                            _ptr]),
    "dbev_spconv_out_table": (_c_int, [_ptr, _c_int, _iptr, _ptr, _c_int, _ptr, _ptr, _ptr, _c_size,
                                       _ptr]),
    "dbev_spconv_pairs_from_table": (_c_int, [_ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr]),
    "dbev_spconv_table_from_pairs": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr,
                                              _ptr]),
    "dbev_spconv_forward": (_c_int, [_ptr, _c_int, _ptr, _c_int, _ptr, _c_int, _c_int, _ptr, _ptr,
                                     _ptr, _c_int, _ptr, _ptr]),
    "dbev_shift_feature_forward": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr]),
    "dbev_shift_feature_backward": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr]),
    "dbev_depth_loss_workspace_bytes": (_c_size, []),
    "dbev_depth_loss_forward": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_float, _c_float, _c_float, _ptr,
                                         _ptr, _c_size, _ptr]),
    "dbev_depth_loss_backward": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_float, _c_float, _c_float, _ptr,
                                          _ptr, _ptr]),
    "dbev_ms_deform_attn_forward": (_c_int, [_ptr] * 5 + [_c_int] * 7 + [_ptr, _ptr]),
    "dbev_ms_deform_attn_backward": (_c_int, [_ptr] * 6 + [_c_int] * 7 + [_ptr, _ptr, _ptr, _ptr]),
    "dbev_center_targets": (_c_int, [_ptr, _c_int, _ptr, _ptr, _c_int, _iptr, _iptr] + [_c_int] * 5
                            + [_c_float] * 6 + [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "dbev_conv2d_tc_forward": (_c_int, [_ptr] + [_c_int] * 4 + [_ptr] + [_c_int] * 5 + [_ptr, _ptr, _c_int, _ptr]
                               + [_c_int] * 8 + [_ptr]),
    "dbev_conv2d_tc_forward_grouped": (_c_int, [_ptr] + [_c_int] * 4 + [_ptr] + [_c_int] * 5 + [_ptr, _ptr, _c_int, _ptr]
                                       + [_c_int] * 9 + [_ptr]),
    "dbev_conv2d_tc_forward_ex": (_c_int, [_ptr] + [_c_int] * 5 + [_ptr] + [_c_int] * 6 + [_ptr, _ptr, _c_int, _ptr]
                                  + [_c_int] * 12 + [_ptr]),
    "dbev_conv_wgrad_tc_workspace_bytes": (_c_size, [_c_int] * 8),
    "dbev_conv_wgrad_tc": (_c_int, [_ptr] + [_c_int] * 5 + [_ptr] + [_c_int] * 8 + [_ptr, _c_int, _ptr, _c_size, _ptr]),
    "dbev_conv2d_tc_dgrad_s2": (_c_int, [_ptr] + [_c_int] * 5 + [_ptr] + [_c_int] * 3 + [_ptr] + [_c_int] * 3 + [_ptr]),
    "dbev_pack_conv_weights": (_c_int, [_ptr] + [_c_int] * 5 + [_ptr, _ptr]),
    "dbev_pack_conv_weights_train": (_c_int, [_ptr] + [_c_int] * 5 + [_ptr, _ptr, _ptr]),
    "dbev_pack_conv_weights_batch": (_c_int, [_ptr, _c_int, _c_int, _ptr]),
    "dbev_channel_stats_workspace_bytes": (_c_size, [_c_ll, _c_int]),
    "dbev_bn_batch_stats": (_c_int, [_ptr, _c_int, _c_ll, _c_int, _ptr, _ptr, _c_float, _c_float, _ptr, _ptr, _ptr, _ptr,
                                     _c_size, _ptr]),
    "dbev_channel_sums": (_c_int, [_ptr, _c_int, _c_ll, _c_int, _ptr, _c_int, _ptr, _c_size, _ptr]),
    "dbev_bn_act_forward": (_c_int, [_ptr, _c_int, _ptr, _ptr, _c_int, _c_ll, _c_int, _c_int, _ptr, _c_int, _ptr, _ptr]),
    "dbev_bn_backward": (_c_int, [_ptr, _c_int, _ptr, _c_int, _ptr, _c_int, _ptr, _c_ll, _c_int, _ptr, _ptr, _c_int, _ptr,
                                  _c_int, _c_int, _ptr, _ptr, _c_size, _ptr]),
    "dbev_relu_mask_backward": (_c_int, [_ptr, _c_int, _ptr, _c_int, _c_ll, _c_int, _ptr, _c_int, _c_int, _ptr, _ptr]),
    "dbev_upsample_bilinear_forward": (_c_int, [_ptr] + [_c_int] * 7 + [_ptr, _c_int, _ptr]),
    "dbev_upsample_bilinear_backward": (_c_int, [_ptr] + [_c_int] * 7 + [_ptr, _c_int, _c_int, _ptr]),
    "dbev_sca_gather_rows": (_c_int, [_ptr, _ptr, _ptr] + [_c_int] * 5 + [_c_ll, _c_ll, _c_ll, _ptr, _ptr]),
    "dbev_sca_reduce_rows": (_c_int, [_ptr, _ptr, _ptr] + [_c_int] * 5 + [_ptr, _ptr]),
    "dbev_hard_pillar_encode": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _fptr, _c_float, _c_float, _ptr, _c_int,
                                         _ptr, _ptr, _c_int, _ptr, _ptr]),
    "dbev_spconv_tc_supported": (_c_int, [_c_int, _c_int, _c_int]),
    "dbev_spconv_pack_weights": (_c_int, [_ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr]),
    "dbev_spconv_forward_tc": (_c_int, [_ptr, _c_int, _ptr, _ptr, _c_int, _ptr, _c_int, _c_int, _ptr,
                                        _ptr, _ptr, _c_int, _ptr, _ptr]),
    "dbev_spconv_dense": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr,
                                   _ptr]),
    "dbev_hard_simple_vfe": (_c_int, [_ptr, _ptr, _c_ll, _c_int, _c_int, _c_int, _ptr, _ptr]),
    "dbev_dynvoxel_coords": (_c_int, [_ptr, _c_int, _c_int, _ptr, _c_int, _fptr, _fptr, _c_int, _ptr,
                                      _ptr]),
    "dbev_dynvoxel_virtual_rows": (_c_int, [_ptr, _c_int, _c_int, _ptr, _ptr]),
    "dbev_dynvoxel_virtual_fix": (_c_int, [_ptr, _ptr, _c_int, _ptr, _ptr]),
    "dbev_affinity_select_workspace_bytes": (_c_size, [_c_int, _c_int]),
    "dbev_affinity_select": (_c_int, [_ptr, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _c_size, _ptr]),
    "dbev_affinity_gather_rows": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr,
                                           _ptr]),
    "dbev_affinity_partial_floats": (_c_size, [_iptr, _c_int]),
    "dbev_affinity_forward": (_c_int, [_ptr, _ptr, _iptr, _c_int, _c_int, _c_int, _c_float, _c_float,
                                       _ptr, _ptr, _ptr]),
    "dbev_affinity_backward": (_c_int, [_ptr, _ptr, _iptr, _c_int, _c_int, _c_int, _c_float,
                                        _c_float, _ptr, _ptr, _ptr]),
    "dbev_affinity_scatter_rows": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr,
                                            _ptr]),
    "dbev_sort_workspace_bytes": (_c_size, [_c_ll]),
    "dbev_sort_keys_iota": (_c_int, [_ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _c_size, _ptr]),
    "dbev_scan_workspace_bytes": (_c_size, [_c_ll]),
    "dbev_exclusive_scan_i32": (_c_int, [_ptr, _ptr, _c_int, _ptr, _ptr, _c_size, _ptr]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "distill_bev_b200: %s is missing - build it with "
            "`python distill-bev_b200/build.py` (nvcc, sm_100a). There is no CPU fallback."
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.dbev_abi_version() != 1:
        raise RuntimeError("distill_bev_b200: ABI version mismatch, rebuild the library")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dbev_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def stream_ptr(device):
    """torch's current stream on `device` as a cudaStream_t (raw query: no Stream object)."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError(
            "%s must be a CUDA tensor: distill_bev_b200 has no CPU path (got device %s)"
            % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("%s must have dtype %s (got %s)" % (name, dtype, t.dtype))
    return t


def h2d_async(host_tensor, device):
    """Small host tensor -> device without stalling the stream: a pageable cudaMemcpyAsync makes the
    driver synchronise the stream first, a pinned source does not (torch's pinned cache keeps the
    staging block alive until the copy has run)."""
    return host_tensor.contiguous().pin_memory().to(device, non_blocking=True)


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def host_f3(vals):
    return (ctypes.c_float * 3)(*[float(v) for v in vals])


def host_i3(vals):
    return (ctypes.c_int * 3)(*[int(v) for v in vals])


def host_floats(vals):
    vals = [float(v) for v in vals]
    return (ctypes.c_float * len(vals))(*vals)


def host_ints(vals):
    vals = [int(v) for v in vals]
    return (ctypes.c_int * len(vals))(*vals)
