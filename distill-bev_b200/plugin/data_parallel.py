"""Data-parallel gradient all-reduce of the trained part of the hot path.

Reference: the training loop wraps the detector in ``MMDistributedDataParallel`` (tools/distributed.py:11-79,
mmdet3d/apis/mmdet_train.py:76-80, launched by tools/dist_train.sh:7-9): torch DDP's reducer averages the gradients
of every student parameter over the ranks with bucketed NCCL all-reduces that overlap the backward pass. The frozen
teacher is held outside ``parameters()`` (bevdet_distill.py:1599-1610) and is not reduced; BatchNorm statistics stay
per-GPU (plain BN in the shipped configs).

``GradientAllReduce`` is that reducer for a CUDA-graph-replayed step: DDP's Python hooks cannot be replayed, so the
gradients live in flat per-bucket buffers (``p.grad`` are views; autograd accumulates into them in place) and each
bucket is reduced with ONE collective - optionally compressed to bf16 like DDP's ``bf16_compress_hook``
(divide by the world size, cast, all-reduce, cast back). Buckets follow the order in which gradients become ready
(reverse registration order, as DDP does), so with ``overlap=True`` a bucket's all-reduce is issued on a side
stream from the post-accumulate hook of its last parameter while the rest of the backward still runs - inside a
CUDA graph capture the NCCL kernels are recorded like any other. Works with any torch.distributed backend (the
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
        self.comm_dtype = comm_dtype
        self.overlap = bool(overlap) and self.world > 1
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("GradientAllReduce: parameters must share one device and dtype")
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
        self.flat, self.comm = [], []
        for bucket in self.buckets:
            n = sum((p.numel() + 3) // 4 * 4 for p in bucket)          # 16-byte aligned views
            flat = torch.zeros(n, dtype=dt, device=dev)
            off = 0
            for p in bucket:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += (p.numel() + 3) // 4 * 4
            self.flat.append(flat)
            self.comm.append(torch.empty(n, dtype=comm_dtype, device=dev) if comm_dtype not in (None, dt) else None)
        self.bytes_per_step = sum((c if c is not None else f).numel() * (c if c is not None else f).element_size()
                                  for f, c in zip(self.flat, self.comm))
        self._pending, self._left, self._hooks = [], [], []
        self._stream = torch.cuda.Stream(dev) if (self.overlap and dev.type == "cuda") else None
        if self.overlap:
            for bi, bucket in enumerate(self.buckets):
                for p in bucket:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
            self.begin()

    # ------------------------------------------------------------------ per step
    def begin(self):
        """Before the backward pass: gradients start from zero (autograd adds into the flat views)."""
        for flat in self.flat:
            flat.zero_()
        self._left = [len(b) for b in self.buckets]
        self._pending = []

    def _make_hook(self, bi):
        def hook(_param):
            self._left[bi] -= 1
            if self._left[bi] == 0:
                self._launch(bi)
        return hook

    def _reduce_bucket(self, bi):
        flat, comm = self.flat[bi], self.comm[bi]
        if self.world == 1:
            return None
        flat.div_(self.world)                      # average, pre-divided like DDP (keeps bf16 in range)
        if comm is not None:
            comm.copy_(flat)
            work = dist.all_reduce(comm, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return work

    def _launch(self, bi):
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat[bi].device))
            with torch.cuda.stream(self._stream):
                work = self._reduce_bucket(bi)
                if work is not None:
                    work.wait()                    # stream-level wait: orders the cast-back after the collective
                if self.comm[bi] is not None:
                    self.flat[bi].copy_(self.comm[bi])
        else:
            self._pending.append((bi, self._reduce_bucket(bi)))

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
            else:
                self._drain()
            return
        for bi in range(len(self.buckets)):
            self._pending.append((bi, self._reduce_bucket(bi)))
        self._drain()

    def _drain(self):
        for bi, work in self._pending:
            if work is not None:
                work.wait()
            if self.comm[bi] is not None:
                self.flat[bi].copy_(self.comm[bi])
        self._pending = []

    def describe(self):
        return {"collective": "all_reduce(SUM) of %d flat gradient bucket(s), pre-divided by the world size"
                              % len(self.buckets),
                "bytes_per_step": int(self.bytes_per_step), "buckets": [int(f.numel()) for f in self.flat],
                "comm_dtype": str(self.comm_dtype or self.flat[0].dtype).replace("torch.", ""),
                "overlap_with_backward": self.overlap, "parameters": len(self.params)}
