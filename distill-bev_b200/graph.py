"""CUDA-graph capture of a hot-path step.

Every entry point of the library is stream-ordered and free of host synchronisation (outputs and
workspaces come from torch's caching allocator, device-side counts stay on the device), so a whole
training-step slice - geometry, plan, lift+splat forward/backward, pillar canvas, adaptation conv
and distillation loss with its backward - can be captured ONCE and replayed: the ~100 launches
of a step are then issued by the GPU front end instead of Python + ctypes (the step is otherwise
bound by CPU issue time, see profiles/r01_step_timeline.txt).

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
