"""CUDA-graph capture of a hot-path step.

The entry points of the training step are stream-ordered and free of host synchronisation (outputs and
workspaces come from torch's caching allocator, device-side counts stay on the device), so the whole step -
geometry, point cells / plan, lift+splat forward/backward, the student BEV encoder (tcgen05 forward / input
gradient / weight gradient, BatchNorm, upsampling), pillar canvas + teacher convs, adaptation conv, distillation
loss with its backward, the NCCL gradient all-reduce and the fused optimizer step - is captured ONCE and replayed:
its ~350 launches are then issued by the GPU front end instead of Python + ctypes (the step is otherwise bound by
CPU issue time, see profiles/r01_step_timeline.txt).

NOT capturable (they return host-side counts, as the reference's API implies): ``Voxelization`` / ``hard_voxelize``
(voxel count as a Python int; ``voxel_layer.hard_voxelize_device`` keeps it on the device),
``dynamic_point_to_voxel_forward`` (output row count), ``affinity.select_rows`` (row offsets), the spconv
rulebook builders (output counts), ``SpatialCrossAttention`` (longest per-camera query list).

Rules for the captured callable: read inputs only from tensors that live across replays (copy the
new batch INTO them before each replay), no `.item()` / `.cpu()`, no pageable host-to-device
copies (host lists such as ground-truth boxes go in as `fgd.PackedBoxes` built on static
buffers).
"""
import torch


class CapturedStep(object):
    """``out = CapturedStep(fn)``; ``out.replay()`` re-runs ``fn``'s kernels and returns the same
    output tensors (overwritten in place by every replay)."""

    def __init__(self, fn, warmup=3, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("CapturedStep needs a CUDA device (distill_bev_b200 has no CPU path)")
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up off the capture: lazy inits, caches, autotune
                for _ in range(max(int(warmup), 1)):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = fn()

    def replay(self):
        self.graph.replay()
        return self.outputs
