"""Data-parallel gradient all-reduce of the trained part of the hot path.

Reference: the training loop wraps the detector in ``MMDistributedDataParallel`` (tools/distributed.py:11-79,
mmdet3d/apis/mmdet_train.py:76-80, launched by tools/dist_train.sh:7-9): torch DDP's reducer averages the gradients
of every student parameter over the ranks with bucketed NCCL all-reduces that overlap the backward pass. The frozen
teacher is held outside ``parameters()`` (bevdet_distill.py:1599-1610) and is not reduced; BatchNorm statistics stay
per-GPU (plain BN in the shipped configs).

``GradientAllReduce`` is that reducer for a CUDA-graph-replayed step (DDP's Python-side reducer cannot be replayed).
Parameters are grouped into buckets in the order their gradients become ready (reverse registration order, as DDP
does). Per bucket and step:

    one multi-tensor copy  gradients -> flat wire buffer (cast to bf16 like DDP's bf16_compress_hook, or fp32)
    ONE all-reduce of the flat buffer (AVG on NCCL; SUM + divide elsewhere)
    one multi-tensor copy  flat wire buffer -> the same gradient tensors

The gradients themselves stay where autograd put them (``p.grad`` is set to None before the backward so the
accumulation step adopts the incoming tensor: no zero-fill, no add). With ``overlap=True`` a bucket is launched on a
side stream from the post-accumulate hook of its last parameter while the rest of the backward still runs; inside a
CUDA graph capture the NCCL kernel is recorded like any other kernel. Works with any torch.distributed backend (the
CPU tests use gloo, world_size 2).
"""
import torch
import torch.distributed as dist


class GradientAllReduce(object):
    def __init__(self, params, world_size=None, process_group=None, bucket_bytes=32 << 20, comm_dtype=None,
                 overlap=False):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradientAllReduce: no trainable parameters")
        self.group = process_group
        self.world = int(world_size if world_size is not None else (dist.get_world_size(process_group)
                                                                    if dist.is_initialized() else 1))
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("GradientAllReduce: parameters must share one device and dtype")
        self.comm_dtype = comm_dtype if comm_dtype is not None else dt
        self.overlap = bool(overlap) and self.world > 1
        self.use_avg = self.world > 1 and dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        # buckets in the order gradients become ready: last registered parameter first
        self.buckets, cur, cur_bytes = [], [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and cur_bytes + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self.flat, self.views = [], []
        for bucket in self.buckets:
            n = sum((p.numel() + 7) // 8 * 8 for p in bucket)          # 16-byte aligned views for bf16 and fp32
            flat = torch.zeros(n, dtype=self.comm_dtype, device=dev)
            views, off = [], 0
            for p in bucket:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += (p.numel() + 7) // 8 * 8
            self.flat.append(flat)
            self.views.append(views)
        self.bytes_per_step = sum(f.numel() * f.element_size() for f in self.flat)
        self._left, self._hooks = [len(b) for b in self.buckets], []
        self._stream = torch.cuda.Stream(dev) if (self.overlap and dev.type == "cuda") else None
        if self.overlap:
            for bi, bucket in enumerate(self.buckets):
                for p in bucket:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    # ------------------------------------------------------------------ per step
    def begin(self):
        """Before the backward pass: drop the old gradients so that autograd adopts the new tensors (no add)."""
        for p in self.params:
            p.grad = None
        self._left = [len(b) for b in self.buckets]

    def _make_hook(self, bi):
        def hook(_param):
            self._left[bi] -= 1
            if self._left[bi] == 0:
                self._launch(bi)
        return hook

    def _reduce_bucket(self, bi):
        grads = [p.grad for p in self.buckets[bi]]
        if any(g is None for g in grads):
            raise RuntimeError("GradientAllReduce: a parameter of bucket %d received no gradient this step" % bi)
        flat, views = self.flat[bi], self.views[bi]
        torch._foreach_copy_(views, grads)                               # gather + cast, one multi-tensor kernel
        if self.use_avg:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            flat.div_(self.world)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        torch._foreach_copy_(grads, views)                               # scatter + cast back, in place

    def _join_producers(self, stream):
        """Weight gradients may have been computed on the conv kernels' side stream (plugin/ops/conv_train.py)."""
        try:
            from .ops import conv_train as ct
        except Exception:  # noqa: BLE001 - CPU-only installs (gloo tests) have no CUDA library
            return
        dev = self.flat[0].device
        if dev.type == "cuda":
            ct.join_side_stream(dev, stream)

    def _launch(self, bi):
        if self.world == 1:
            return
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat[bi].device))
            self._join_producers(self._stream)
            with torch.cuda.stream(self._stream):
                self._reduce_bucket(bi)
        else:
            self._join_producers(None)
            self._reduce_bucket(bi)

    def finish(self):
        """After the backward pass: every ``p.grad`` holds the rank-averaged gradient when this returns (stream-ordered)."""
        if self.world == 1:
            return
        if self.overlap:
            if any(n != 0 for n in self._left):
                raise RuntimeError("GradientAllReduce: %d bucket(s) saw no gradient for some parameter this step"
                                   % sum(1 for n in self._left if n != 0))
            if self._stream is not None:
                torch.cuda.current_stream(self.flat[0].device).wait_stream(self._stream)
            return
        self._join_producers(None)
        for bi in range(len(self.buckets)):
            self._reduce_bucket(bi)

    def describe(self):
        return {"collective": "%d x all_reduce(%s) of a flat gradient bucket, averaged over the ranks"
                              % (len(self.buckets), "AVG" if self.use_avg else "SUM, pre-divided"),
                "bytes_per_step": int(self.bytes_per_step), "buckets": [int(f.numel()) for f in self.flat],
                "comm_dtype": str(self.comm_dtype).replace("torch.", ""),
                "overlap_with_backward": self.overlap, "parameters": len(self.params)}
