"""distill-bev_b200: B200-native (sm_100a) DistillBEV hot path behind the
reference's mmdet3d operator names. See DESIGN.md and INTEGRATION.md.

Import as ``distill_bev_b200`` (alias module at the repository root).
"""
from . import _lib  # noqa: F401
from .plugin.ops.bev_pool import (BevPlan, GridSpec, PointCells, QuickCumsumCuda, bev_plan_from_coords,  # noqa: F401
                                  bev_point_cells,
                                  bev_plan_from_geom, bev_pool, bev_pool_ext, bev_pool_gather,
                                  lift_splat, transpose_batched, voxel_pooling)

from .plugin.ops.voxel import (DynamicScatter, Voxelization, dynamic_scatter, voxel_layer,  # noqa: F401
                               voxelization)

from .plugin.pillars import DynamicPillarFeatureNet, PFNLayer, PillarFeatureNet, PointPillarsScatter, pillar_canvas  # noqa: F401
from .plugin.view_transformer import (ModulatedDeformConv2dPack, SELikeModule, ViewTransformerLiftSplatShoot,  # noqa: F401
                                      ViewTransformerLSSBEVDepth, lss_geometry)
from .plugin.distill import fgd  # noqa: F401
from .plugin.distill import affinity, bevformer, detector  # noqa: F401
from .plugin.distill.adaptation import Conv1x1Adaptation, conv1x1  # noqa: F401
from .plugin.ops import spconv  # noqa: F401
from .plugin.ops.ms_deform_attn import (MultiScaleDeformableAttnFunction,  # noqa: F401
                                        MultiScaleDeformableAttnFunction_fp32, multi_scale_deformable_attn)
from .plugin.sparse_teacher import DynamicVoxelEncoder, HardSimpleVFE, SparseEncoder  # noqa: F401
from .plugin.dense_teacher import SECOND, SECONDFPN  # noqa: F401
from .plugin.student_convs import Conv2dTC, conv2d_tc, conv2d_tc_supported, convert_convs  # noqa: F401
from .plugin.bev_encoder import BasicBlock, Bottleneck, FPN_LSS, ResNetForBEVDet, conv_bn_act, upsample_cat  # noqa: F401
from .plugin.ops import conv_train  # noqa: F401
from .plugin import bev_encoder  # noqa: F401
from .plugin.bevformer_attention import MSDeformableAttention3D, SpatialCrossAttention  # noqa: F401
from .plugin.bevdepth import get_depth_loss, shift_feature  # noqa: F401
from .plugin.center_targets import CenterHeadTargets  # noqa: F401
from .graph import CapturedStep  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
