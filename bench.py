"""bench.py — DistillBEV hot-path training step on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one TRAINING pass of the hot path over one synthetic nuScenes-shaped batch per GPU
(BASELINE.json configs[1]: CenterPoint -> BEVDepth-R50 distillation, per-GPU batch 8, 2 frames, 6 cams, D=59,
16x44 frustum, C=64 -> 128x128 BEV; 30k-point LiDAR; head position 256 -> 384 ch), one autograd chain:
  A  student view transform: get_geometry -> point cells -> fused lift+splat (16 sample-frames), frames concatenated
  S  student BEV encoder, TRAINING mode: ResNetForBEVDet (3 x 2 BasicBlocks, 128 -> 128/256/512) + FPN_LSS
     (-> 256 ch @ 128x128) on the tcgen05 conv kernels (forward, input gradient, weight gradient), batch-stat BatchNorm
  B  frozen teacher LiDAR path (side stream): voxelize -> DynamicPillarFeatureNet -> PointPillarsScatter ->
     SECOND + SECONDFPN (tcgen05 conv+BN+ReLU) -> teacher BEV feature [8,384,128,128]
  C  head-position distillation loss: fg / fp masks from GT boxes and device-rasterised heat maps, 1x1 adaptation
     conv, fused fgd loss
  backward of C -> S -> A, then (N > 1) the data-parallel gradient all-reduce (NCCL, flat bf16 buckets), then a fused
  AdamW step on the student encoder + adaptation parameters.
The image backbone / depth net / CenterHead (mmdet / mmcv third-party modules outside SURVEY §8) are not in the step.

`value` times the step with every input resident in HBM; `e2e` times the same step when the host-originated inputs
(calibration, LiDAR points, GT boxes, labels) start in pinned host memory and the loss scalars are read back, every
step. `roofline` is measured live for the kernel class that dominates the step (conv3x3_halo_kernel<256>, tensor
bound); `roofline_bev_pool` is BASELINE.json's "bev_pool HBM GB/s vs roofline". `--impl reference` / `cpu_baseline`
time the CPU implementation of the same stages (oracle/ numpy + C; the conv stacks through torch.nn on the CPU, which
IS the reference's implementation of those rows).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "distill_hotpath_samples_per_sec"
UNIT = "samples/s"
BATCH = 8            # per-GPU batch of the shipped config (...r50.py:308)
FRAMES = 2           # BEVDepth4D: current + adjacent frame
N_CAMS, D, FH, FW, C_TRANS, BEV = 6, 59, 16, 44, 64, 128
N_POINTS = 30000
C_STUDENT, C_TEACHER = 256, 384
WORKLOAD = ("distill-train-hotpath-v3 (configs[1], B=8/GPU): lift+splat (2 frames x 6 cams, D=59, 16x44, C=64 -> 128x128, "
            "sort-free splat) -> student BEV encoder in TRAINING mode (ResNetForBEVDet 128->128/256/512 + FPN_LSS -> 256 ch, "
            "tcgen05 TF32 fwd / dgrad / wgrad, batch-stat BN) -> 1x1 adaptation + fgd distill loss at head "
            "(256->384 ch, fg+fp masks) against the frozen LiDAR teacher run end to end in the step (voxelize/pillar-encode/"
            "scatter 8 x 30k pts -> SECOND + SECONDFPN on tcgen05 -> [8,384,128,128]); backward through loss, encoder and "
            "lift+splat; gradient all-reduce when N > 1; fused AdamW step. Image backbone / depth net / CenterHead "
            "(third-party mmdet modules, outside SURVEY 8) not in the step")
ENC_CHANNELS, ENC_OUT = [128, 256, 512], 256
PILLAR_VS, PILLAR_RANGE = [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
CENTER_TASKS = [dict(num_class=1, class_names=['car']), dict(num_class=2, class_names=['truck', 'construction_vehicle']),
                dict(num_class=2, class_names=['bus', 'trailer']), dict(num_class=1, class_names=['barrier']),
                dict(num_class=2, class_names=['motorcycle', 'bicycle']),
                dict(num_class=2, class_names=['pedestrian', 'traffic_cone'])]
DISTILL_PARAMS = dict(
    spatial_t=0.5, spatial_student_ratio=1.0, channel_t=0.5, fg_feat_loss_weights=[6e-3],
    bg_feat_loss_weights=[4e-2], channel_loss_weights=[0.25], spatial_loss_weights=[2.5e-3],
    spatial_attentions=["teacher_student"], transpose_mask=False, foreground_mask="gt",
    background_mask="logical_not", scale_mask="combine_gt", spatial_mask=True, channel_mask=False,
    output_threshold=0.1, groundtruth_threshold=None, fp_as_foreground=["teacher"], fp_weight=6e-2,
    fp_epoch=0, fp_scale_mode="average")
TRAIN_CFG = dict(grid_size=[1024, 1024, 40], point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0],
                 voxel_size=[0.1, 0.1, 0.2])


# ----------------------------------------------------------------------------- helpers

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def aggregate_step_time(ms_local, world, dist=None, device=None):
    """MAX over ranks of the per-rank device time (the contract's max-over-ranks rule) and the
    whole-job throughput: every rank processes its own BATCH samples per step (weak scaling, no
    data-path collective), so value = BATCH * world / max_ms. Works with any backend (tested with
    gloo on CPU, world_size 2)."""
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def samples_per_sec(total_ms, steps, world):
    return BATCH * world / (total_ms / steps * 1e-3)


def rank_seed(rank):
    """Each rank draws its own synthetic batch (independent samples, like a DistributedSampler)."""
    return 1000 + rank


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# ----------------------------------------------------------------------------- our arm

def encoder_flops(batch):
    """Forward FLOPs of the student BEV encoder (2 per multiply-add); training = 3x (fwd + dgrad + wgrad)."""
    f, cin, hw = 0.0, 2 * C_TRANS, BEV
    for c in ENC_CHANNELS:
        hw //= 2
        f += 2.0 * batch * hw * hw * 9 * (2 * cin * c + 3 * c * c)      # conv1 + downsample (stride 2), 3 more 3x3
        cin = c
    h1 = BEV // 2
    f += 2.0 * batch * h1 * h1 * 9 * ((ENC_CHANNELS[0] + ENC_CHANNELS[2]) * 2 * ENC_OUT + 4 * ENC_OUT * ENC_OUT)
    f += 2.0 * batch * BEV * BEV * (9 * 2 * ENC_OUT * ENC_OUT + ENC_OUT * ENC_OUT)
    return f


class CudnnEncoder(object):
    """The same encoder from plain torch modules (what the reference runs: cuDNN TF32 + ATen), for the side-by-side key."""

    @staticmethod
    def build(device):
        import torch
        import torch.nn as nn

        def cbr(cin, cout, k=3, s=1, p=1):
            return nn.Sequential(nn.Conv2d(cin, cout, k, s, p, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

        class Basic(nn.Module):
            def __init__(self, cin, cout, stride=1):
                super().__init__()
                self.c1 = cbr(cin, cout, 3, stride, 1)
                self.c2 = nn.Sequential(nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout))
                self.down = nn.Conv2d(cin, cout, 3, stride, 1) if stride != 1 else None

            def forward(self, x):
                idt = x if self.down is None else self.down(x)
                return torch.relu(self.c2(self.c1(x)) + idt)

        class Encoder(nn.Module):
            def __init__(self):
                super().__init__()
                layers, cin = [], 2 * C_TRANS
                for cout in ENC_CHANNELS:
                    layers.append(nn.Sequential(Basic(cin, cout, 2), Basic(cout, cout)))
                    cin = cout
                self.layers = nn.ModuleList(layers)
                self.up = nn.Upsample(scale_factor=4, mode="bilinear", align_corners=True)
                self.conv = nn.Sequential(cbr(ENC_CHANNELS[0] + ENC_CHANNELS[2], 2 * ENC_OUT), cbr(2 * ENC_OUT, 2 * ENC_OUT))
                self.up2 = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                                         cbr(2 * ENC_OUT, ENC_OUT), nn.Conv2d(ENC_OUT, ENC_OUT, 1))

            def forward(self, x):
                feats = []
                for l in self.layers:
                    x = l(x)
                    feats.append(x)
                return self.up2(self.conv(torch.cat([feats[0], self.up(feats[2])], 1)))

        torch.manual_seed(0)
        return Encoder().to(device).train().to(memory_format=torch.channels_last)


class HotPath(object):
    """Synthetic batch + the training step, built on the public plugin API (distill_bev_b200)."""

    def __init__(self, device, seed, world=1, allreduce="overlap", comm_dtype="bf16", encoder="tcgen05"):
        import torch
        import distill_bev_b200 as dbev
        from distill_bev_b200 import synthetic
        self.torch, self.dbev, self.dev, self.world = torch, dbev, device, world
        g = torch.Generator().manual_seed(seed)
        nf = BATCH * FRAMES
        # --- host-originated inputs (pinned) -------------------------------------------------
        calib = synthetic.make_calibration(nf, N_CAMS, seed=seed)
        self.h_calib = [torch.from_numpy(a).pin_memory() for a in calib]
        clouds = synthetic.make_lidar(BATCH, N_POINTS, seed=seed)
        self.h_points = [torch.from_numpy(c).pin_memory() for c in clouds]
        gt = synthetic.make_gt_boxes(BATCH, seed=seed)
        self.boxes = [torch.from_numpy(b) for b, _ in gt]
        self.labels = [torch.from_numpy(np.asarray(l)).to(torch.int32) for _, l in gt]
        self.h_labels = torch.cat(self.labels).contiguous().pin_memory()
        # CenterHead targets (GT heat maps of add_fp_as_fg) are rasterised on the device from the boxes
        self.targets = dbev.CenterHeadTargets(CENTER_TASKS, dict(TRAIN_CFG, out_size_factor=8, dense_reg=1,
                                                                  gaussian_overlap=0.1, max_objs=500, min_radius=2))
        # --- device-resident activations produced by the image branch (outside SURVEY 8) ------
        self.depth = torch.randn(nf * N_CAMS, D, FH, FW, generator=g).softmax(1).to(device).requires_grad_(True)
        self.feat = torch.randn(nf * N_CAMS, C_TRANS, FH, FW, generator=g).to(device).requires_grad_(True)
        self.sorted_splat = False
        self.teacher_logit = (torch.randn(BATCH, 10, BEV, BEV, generator=g) * 1.5 - 3.0).to(device)
        # --- modules ---------------------------------------------------------------------------
        torch.manual_seed(0)          # identical initial weights on every rank, like DDP's broadcast at construction
        self.vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=32,
                                                     numC_Trans=C_TRANS).to(device)
        self.enc = dbev.DynamicPillarFeatureNet(in_channels=5, feat_channels=(64,), voxel_size=PILLAR_VS,
                                                point_cloud_range=PILLAR_RANGE).to(device).eval()
        self.scat = dbev.PointPillarsScatter(64, [512, 512], channels_last=True)
        # teacher BEV backbone + neck (centerpoint_02pillar_second_secfpn: SECOND 64 -> 64/128/256, FPN 3 x 128)
        self.second = dbev.SECOND(in_channels=64, out_channels=[64, 128, 256], layer_nums=[3, 5, 5],
                                  layer_strides=[2, 2, 2]).to(device).eval()
        self.secfpn = dbev.SECONDFPN(in_channels=[64, 128, 256], out_channels=[128, 128, 128],
                                     upsample_strides=[0.5, 1, 2]).to(device).eval()
        # student BEV encoder (img_bev_encoder_backbone / img_bev_encoder_neck of ...bevdepth4d_r50.py:122-126)
        self.encoder_kind = encoder
        if encoder == "cudnn":
            self.student_net = CudnnEncoder.build(device)
        else:
            self.backbone = dbev.ResNetForBEVDet(numC_input=2 * C_TRANS, num_channels=ENC_CHANNELS).to(device).train()
            self.neck = dbev.FPN_LSS(in_channels=ENC_CHANNELS[0] + ENC_CHANNELS[2], out_channels=ENC_OUT).to(device).train()
            self.student_net = torch.nn.ModuleList([self.backbone, self.neck])
        from distill_bev_b200.plugin.distill.adaptation import Conv1x1Adaptation
        self.adapt = Conv1x1Adaptation(C_STUDENT, C_TEACHER).to(device)            # '1x1conv' adaptation
        self.spatial = torch.nn.Conv2d(1, 1, 3, padding=1).to(device)            # spatial_wise_adaptations
        self.trainable = [p for m in (self.student_net, self.adapt, self.spatial) for p in m.parameters()]
        self.n_params = sum(p.numel() for p in self.trainable)
        self.optim = torch.optim.AdamW(self.trainable, lr=2e-4, weight_decay=0.01, fused=True, capturable=True)
        self.reducer, self.allreduce, self.comm_dtype = None, allreduce, comm_dtype
        self.set_allreduce(allreduce)
        # weight gradients and the per-step weight packing run beside the critical path (conv_train side stream)
        self.overlap_side = encoder != "cudnn" and os.environ.get("DBEV_BENCH_SIDE_STREAM", "1") != "0"
        dbev.conv_train.set_side_stream(self.overlap_side)
        self.d_calib = [t.to(device) for t in self.h_calib]
        self.d_points = [t.to(device) for t in self.h_points]
        counts = [int(b.shape[0]) for b in self.boxes]
        self.max_boxes = max(counts + [1])
        self.h_boxes = torch.cat([b.float() for b in self.boxes], 0).contiguous().pin_memory()
        self.h_box_offs = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32).pin_memory()
        self.d_boxes, self.d_box_offs = self.h_boxes.to(device), self.h_box_offs.to(device)
        self.d_labels = self.h_labels.to(device)
        self.captured, self.split = None, False
        self.side = [torch.cuda.Stream(device)]
        self.opt_stream = torch.cuda.Stream(device)
        self.h2d_bytes = (sum(t.numel() * 4 for t in self.h_calib) + sum(t.numel() * 4 for t in self.h_points)
                          + self.h_labels.numel() * 4 + self.h_box_offs.numel() * 4
                          + sum(b.numel() * 4 for b in self.boxes))
        self.d2h_bytes = 5 * 4

    def set_allreduce(self, mode):
        """(Re)build the gradient reducer: 'overlap' = per-bucket all-reduce launched from gradient hooks on a side
        stream (recorded into the CUDA graph), 'after' = all buckets after the backward (eager NCCL between two graphs),
        'none' = no collective (replicas; not data-parallel training)."""
        torch = self.torch
        if self.reducer is not None:
            for h in self.reducer._hooks:
                h.remove()
        self.reducer, self.allreduce = None, mode
        if self.world > 1 and mode != "none":
            from distill_bev_b200.plugin.data_parallel import GradientAllReduce
            self.reducer = GradientAllReduce(self.trainable, self.world,
                                             comm_dtype=torch.bfloat16 if self.comm_dtype == "bf16" else None,
                                             overlap=(mode == "overlap"))

    def encode(self, bev):
        if self.encoder_kind == "cudnn":
            return self.student_net(bev)
        return self.neck(self.backbone(bev))

    def _forward_backward(self, calib, points, labels, boxes):
        """Forward and backward of one batch (public plugin API only). The frozen teacher (B) is independent of the
        student chain (A -> S) until the loss, so it runs on a side stream and the loss waits for its event."""
        torch, dbev = self.torch, self.dbev
        main = torch.cuda.current_stream(self.dev)
        if self.encoder_kind != "cudnn":
            dbev.conv_train.set_side_stream(self.overlap_side)
            dbev.bev_encoder.prepack(self.student_net)     # this step's weights -> TMA matrices, beside lift+splat / teacher
        self.side[0].wait_stream(main)
        with torch.cuda.stream(self.side[0]), torch.no_grad():
            canvas = dbev.pillar_canvas(points, self.enc, self.scat)
            teacher = self.secfpn(self.second(canvas))[0].contiguous()      # NCHW for the loss kernels
            teacher_ready = torch.cuda.Event()
            teacher_ready.record(self.side[0])
        # A: student view transform (geometry changes every step: augmentation)
        geom = self.vt.get_geometry(*calib)
        if self.sorted_splat:
            bev = dbev.lift_splat(self.depth, self.feat, self.vt.make_plan(geom, BATCH * FRAMES))     # [B*2, 64, 128, 128]
            # the two frames of a sample concatenated along the channels (bevdet.py:300-320), channels-last memory
            bev = bev.permute(0, 2, 3, 1).reshape(BATCH, FRAMES, BEV, BEV, C_TRANS).permute(0, 2, 3, 1, 4) \
                     .reshape(BATCH, BEV, BEV, FRAMES * C_TRANS).permute(0, 3, 1, 2)
        else:
            # sort-free splat with the frame index numbered last: the result IS the channel concat of the two frames
            bev = dbev.lift_splat(self.depth, self.feat, self.vt.make_cells(geom, BATCH * FRAMES, frames=FRAMES))
        # the encoder's last parameter gradient exists once d loss / d bev does (before the splat's backward starts):
        # the optimizer step waits for this event only and overlaps the rest of the backward chain (the splat's
        # backward here; the image-view network's backward in a full detector)
        self._enc_grads_done = torch.cuda.Event()
        bev.register_hook(lambda g: self._enc_grads_done.record(torch.cuda.current_stream(self.dev)))
        # S: student BEV encoder (training mode)
        s_feat = self.encode(bev)
        # C: head-position distillation loss (1x1 channel adaptation inside, as in the reference)
        self.targets.get_targets(boxes, labels, device=self.dev)
        gt_hm = self.targets.last_heatmap
        losses = dbev.fgd.fgd_distill_loss(
            teacher, s_feat, boxes, DISTILL_PARAMS, TRAIN_CFG, channel_adaptation=self.adapt, teacher_ready=teacher_ready,
            spatial_adaptation=self.spatial, heatmaps=gt_hm, teacher_heatmaps=self.teacher_logit, epoch=1)
        total = losses["kd_fg_feat_loss"] + losses["kd_bg_feat_loss"] + losses["kd_spatial_loss"] \
            + losses["kd_fp_bg_feat_loss"]
        if self.reducer is not None:
            self.reducer.begin()
        total.backward()
        loss_vec = torch.stack([losses[k] for k in sorted(losses)]).detach()
        dbev.conv_train.join_side_stream(self.dev)            # weight gradients computed beside the backward chain
        main.wait_stream(self.side[0])
        for t in (canvas, teacher):
            t.record_stream(main)
        for p in (self.depth, self.feat):
            p.grad = None
        return loss_vec, canvas

    def _update(self, overlap_event=None, join=True):
        """AdamW over the trained parameters. With ``overlap_event`` (recorded when the last parameter gradient of the
        main stream exists) the step runs on its own stream beside the tail of the backward chain. ``join=False``: the
        caller has already joined the weight-gradient stream (the optimizer graph of the split mode is captured on its own:
        a wait on the uncaptured side stream would invalidate that capture)."""
        torch = self.torch
        if overlap_event is None or not self.overlap_side:
            if join:
                self.dbev.conv_train.join_side_stream(self.dev)   # weight gradients computed beside the backward chain
            self.optim.step()
        else:
            main = torch.cuda.current_stream(self.dev)
            self.opt_stream.wait_event(overlap_event)
            self.dbev.conv_train.join_side_stream(self.dev, self.opt_stream)
            with torch.cuda.stream(self.opt_stream):
                if self.reducer is not None:
                    self.reducer.finish()                     # the optimizer stream waits for the all-reduce stream
                self.optim.step()
            main.wait_stream(self.opt_stream)
        if self.reducer is None:
            self.optim.zero_grad(set_to_none=True)

    def _compute(self, calib, points, labels, boxes):
        """One whole training step issued in one go (used eagerly and as the single captured graph when the gradient
        all-reduce is absent or captured with it)."""
        out = self._forward_backward(calib, points, labels, boxes)
        ev = self._enc_grads_done if self.overlap_side else None
        if ev is None and self.reducer is not None:
            self.reducer.finish()
        self._update(ev)
        return out

    def _device_inputs(self):
        d = dict(calib=[t.to(self.dev) for t in self.h_calib], points=[t.to(self.dev) for t in self.h_points],
                 labels=self.h_labels.to(self.dev), boxes=self.h_boxes.to(self.dev),
                 box_offs=self.h_box_offs.to(self.dev))
        d["packed"] = self.dbev.fgd.PackedBoxes(d["boxes"], d["box_offs"], self.max_boxes)
        return d

    def _copy_batch(self, d):
        """This step's inputs: pinned host memory -> the static tensors a captured step reads."""
        for dst, h in zip(d["calib"], self.h_calib):
            dst.copy_(h, non_blocking=True)
        for dst, h in zip(d["points"], self.h_points):
            dst.copy_(h, non_blocking=True)
        d["labels"].copy_(self.h_labels, non_blocking=True)
        d["boxes"].copy_(self.h_boxes, non_blocking=True)
        d["box_offs"].copy_(self.h_box_offs, non_blocking=True)

    def enable_graph(self):
        """Capture the step once per input set; afterwards step() copies the new batch into the static input tensors
        (e2e) and replays. Two input sets + two graphs let the e2e path copy batch i+1 on a side stream while step i
        runs (what a prefetching data loader does). With allreduce='after' the NCCL all-reduce is issued eagerly
        between a captured forward+backward graph and a captured optimizer graph; with 'overlap' / 'none' or N = 1
        the whole step is one graph (NCCL kernels recorded into it)."""
        torch, dbev = self.torch, self.dbev
        self.sets = [dict(calib=self.d_calib, points=self.d_points, labels=self.d_labels, boxes=self.d_boxes,
                          box_offs=self.d_box_offs,
                          packed=dbev.fgd.PackedBoxes(self.d_boxes, self.d_box_offs, self.max_boxes)),
                     self._device_inputs()]
        self.split = self.reducer is not None and self.allreduce == "after"
        if self.split:
            self.graphs = [dbev.CapturedStep(
                (lambda d=d: self._forward_backward(d["calib"], d["points"], d["labels"], d["packed"])), warmup=3,
                device=self.dev) for d in self.sets]
            for _ in range(2):            # optimizer state must exist before its capture
                self.reducer.finish()
                self._update()
            # _forward_backward() ends with the join of the weight-gradient stream, so the optimizer graph needs none
            self.update_graph = dbev.CapturedStep(lambda: (self._update(join=False), self.trainable[0])[1], warmup=1,
                                                  device=self.dev)
        else:
            self.graphs = [dbev.CapturedStep(
                (lambda d=d: self._compute(d["calib"], d["points"], d["labels"], d["packed"])), warmup=3,
                device=self.dev) for d in self.sets]
        self.captured = self.graphs[0]
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]      # batch landed in set k
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]    # graph k finished reading set k
        self.e2e_idx, self.prefetched = 0, False

    def _replay(self, k):
        out = self.graphs[k].replay()
        if self.split:
            self.reducer.finish()
            self.update_graph.replay()
        return out

    def step(self, e2e):
        torch = self.torch
        if self.captured is not None:
            if not e2e:
                loss_vec, canvas = self._replay(0)
                return loss_vec, canvas
            main = torch.cuda.current_stream(self.dev)
            cur, nxt = self.e2e_idx % 2, (self.e2e_idx + 1) % 2
            if not self.prefetched:          # very first e2e step: nothing was prefetched for it
                with torch.cuda.stream(self.copy_stream):
                    self._copy_batch(self.sets[cur])
                    self.copied[cur].record(self.copy_stream)
            main.wait_event(self.copied[cur])
            loss_vec, canvas = self._replay(cur)
            self.consumed[cur].record(main)
            # the NEXT step's batch: H2D on the copy stream while this step computes
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.consumed[nxt])
                self._copy_batch(self.sets[nxt])
                self.copied[nxt].record(self.copy_stream)
            self.prefetched = True
            self.e2e_idx += 1
            return loss_vec.cpu(), canvas     # device -> host read of this step's losses (synchronises)
        if e2e:
            calib = [t.to(self.dev, non_blocking=True) for t in self.h_calib]
            points = [t.to(self.dev, non_blocking=True) for t in self.h_points]
            labels = self.h_labels.to(self.dev, non_blocking=True)
        else:
            calib, points, labels = self.d_calib, self.d_points, self.d_labels
        loss_vec, canvas = self._compute(calib, points, labels, self.boxes)
        return (loss_vec.cpu() if e2e else loss_vec), canvas


def count_our_kernels(hp):
    """Kernel launches of THIS library in one step (names in namespace dbev::), via CUPTI."""
    torch = hp.torch
    captured, hp.captured = hp.captured, None   # count the eager launch sequence (same kernels)
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            hp.step(False)
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        ours = [n for n in names if "dbev::" in n]
        return len(ours), len(names)
    except Exception as exc:  # CUPTI unavailable: static count of the launch sequence
        sys.stderr.write("profiler unavailable (%s); using the static launch count\n" % exc)
        return 45, None
    finally:
        hp.captured = captured


def sparse_teacher_probe(device, batch=2, n_points=240000):
    """Not part of the timed step (configs[1] uses the pillar teacher): the sparse LiDAR teacher of
    configs[3] — hard voxelize (0.064 m voxels, 41 x 1600 x 1600 grid) + HardSimpleVFE + the LidarFormer
    SparseEncoder (20 sparse convs: lane-group kernels for the narrow layers, tcgen05 3xTF32 implicit
    GEMM for C >= 32) -> dense [B, 256, 200, 200], timed with CUDA events (median of 5)."""
    import torch
    import distill_bev_b200 as dbev
    from distill_bev_b200 import synthetic
    vox = dbev.Voxelization([0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 10, (90000, 120000)).eval()
    vfe = dbev.HardSimpleVFE(5)
    enc = dbev.SparseEncoder(
        in_channels=5, sparse_shape=[41, 1600, 1600], output_channels=128,
        encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
        encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type="basicblock").to(device).eval()
    clouds = [torch.from_numpy(c).to(device) for c in synthetic.make_lidar_scene(batch, n_points, seed=3)]

    def front():
        feats, coors = [], []
        for b, pts in enumerate(clouds):
            v, c, n = vox(pts)
            feats.append(vfe(v, n, c))
            coors.append(torch.nn.functional.pad(c, (1, 0), value=b))
        return torch.cat(feats), torch.cat(coors).contiguous()

    def timed_ms(fn, iters=5):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            r = fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2], r

    t_front, (feats, coors) = timed_ms(front)
    t_enc, out = timed_ms(lambda: enc(feats, coors, batch))
    return {"workload": "LidarFormer sparse teacher front end: %d x %d-point clouds -> %d voxels -> %s"
                        % (batch, n_points, feats.shape[0], list(out.shape)),
            "voxelize_vfe_ms": round(t_front, 3), "sparse_encoder_ms": round(t_enc, 3),
            "samples_per_sec": round(batch / ((t_front + t_enc) * 1e-3), 1),
            "note": "includes the host read-backs of voxel / output counts the reference API implies"}


def sparse_conv_roofline(device, batch=2, n_points=240000):
    """`roofline` of the configs[3] line: the kernel class with the largest share of the sparse encoder
    (profiles/r01_spconv.json), a 128 -> 128 submanifold 3x3x3 conv on the tcgen05 3xTF32 kernel, timed alone (CUDA
    events, median of 7) on the voxel set of the encoder's LAST stage (the step's clouds, three strided convs deep). Tensor-bound; useful flops = 2 * pairs * 128 * 128 where a
    pair is a present (output voxel, neighbour) couple; peak = measured TF32 rate / 3 (three TF32 MMAs per fp32-equivalent
    product)."""
    import torch
    import distill_bev_b200 as dbev
    from distill_bev_b200 import synthetic
    from distill_bev_b200.plugin.ops import spconv as sp
    vox = dbev.Voxelization([0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 10, (90000, 120000)).eval()
    coors = []
    for b, pts in enumerate(synthetic.make_lidar_scene(batch, n_points, seed=3)):
        _, c, _ = vox(torch.from_numpy(pts).to(device))
        coors.append(torch.nn.functional.pad(c, (1, 0), value=b))
    coors = torch.cat(coors).contiguous()
    # the voxel set of the encoder's last stage: three strided SparseConv3d (k3 s2; paddings 1, 1, (0, 1, 1)) deep
    shape = [41, 1600, 1600]
    for pad in (1, 1, (0, 1, 1)):
        down = sp.build_rulebook(coors, batch, shape, 3, 2, pad, 1, False)
        coors, shape = down.out_indices.contiguous(), list(down.out_shape)
    rb = sp.build_rulebook(coors, batch, shape, 3, 1, 1, 1, True)
    feats = torch.randn(coors.shape[0], 128, device=device)
    w = torch.randn(3, 3, 3, 128, 128, device=device) * 0.05
    for _ in range(2):
        sp.conv_table(feats, w, rb.nbr, rb.n_out)
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        sp.conv_table(feats, w, rb.nbr, rb.n_out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    pairs = int((rb.nbr >= 0).sum().item())
    flops = 2.0 * pairs * 128 * 128
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak, how = float(json.load(open(path))["bf16_tflops"]) / 6.0, "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 rate) / 3 (3xTF32)"
    else:
        peak, how = 1125.0 / 3.0, "fallback (B200_PROFILING.md dense bf16 / 2 / 3)"
    ach = flops / ms / 1e9
    return {"kernel": "dbev::sp_conv_tc_kernel<128, 27, 4> (SubM 3x3x3 128 -> 128, A operand in tensor memory)", "bound": "tensor",
            "achieved": round(ach, 1), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": None,
            "peak_source": how, "launch_ms": round(ms, 4), "voxels": int(coors.shape[0]), "pairs": pairs, "grid": shape,
            "note": "useful fp32-equivalent flops of present neighbour pairs only (a 27 x 128 tile slot is empty for absent "
                    "neighbours: ~56 % occupancy at LiDAR sparsity); ncu --set full of this kernel (B = 4): profiles/r01_spconv.json (tensor pipe "
                    "52 % of active cycles, DRAM 2 %)"}


def cudnn_encoder_probe(device, seed, steps=10):
    """Side-by-side evidence, NOT the headline: the identical training step with the student BEV encoder built from
    plain torch modules (nn.Conv2d / BatchNorm2d / ReLU / Upsample: cuDNN TF32 + ATen, channels_last) - what the
    reference runs for row S1 - everything else unchanged (our lift+splat, teacher, loss, fused AdamW), same CUDA
    graph capture."""
    import torch
    hp2 = HotPath(device, seed=seed, encoder="cudnn")
    note = "graph"
    try:
        hp2.enable_graph()
    except Exception as exc:  # noqa: BLE001
        hp2.captured, note = None, "eager (%s)" % str(exc).splitlines()[0][:80]
    for _ in range(3):
        hp2.step(False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(steps):
        loss_vec, _ = hp2.step(False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    del hp2
    torch.cuda.empty_cache()
    return {"workload": "the same step with the student BEV encoder through torch modules (cuDNN TF32 channels_last + ATen "
                        "BatchNorm / ReLU / Upsample + autograd), " + note,
            "ms_per_step": round(ms, 3), "samples_per_sec": round(BATCH / (ms * 1e-3), 1),
            "loss_finite": bool(torch.isfinite(loss_vec).all().item())}


def student_conv_roofline(hp):
    """`roofline` of the kernel class with the largest share of the step (profiles/r02_launches_bench_step.json):
    conv3x3_halo_kernel<256>, timed alone with CUDA events on its stream on the step's largest layer (FPN_LSS
    512 -> 256 @ 128x128, B=8: forward; the input gradients of the 3x3 / stride-1 layers run the same kernel).
    peak = half of the measured bf16 rate (kind::tf32 issues at half the bf16 rate on tcgen05). Activations
    (268 MB in, 134 MB out) exceed L2."""
    import torch
    from distill_bev_b200 import conv_train as ct
    dev = hp.dev
    x = torch.randn(BATCH, BEV, BEV, 2 * ENC_OUT, device=dev)
    w = torch.randn(ENC_OUT, 2 * ENC_OUT, 3, 3, device=dev) * 0.05
    wf = ct.pack_weights(w, 0)
    out = torch.empty(BATCH, BEV, BEV, ENC_OUT, device=dev)
    for _ in range(3):
        ct.conv_forward(x, wf, ENC_OUT, 3, 3, 1, 1, out=out)
    iters = 20
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        ct.conv_forward(x, wf, ENC_OUT, 3, 3, 1, 1, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    flops = 2.0 * BATCH * BEV * BEV * 9 * 2 * ENC_OUT * ENC_OUT
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak, how = float(json.load(open(path))["bf16_tflops"]) / 2.0, "MEASURED_PEAKS.json bf16_tflops (burst) / 2 = TF32 rate"
    else:
        peak, how = 1100.0, "fallback: nominal dense TF32 1.1 PFLOP/s (B200_PROFILING.md)"
    ach = flops / (ms * 1e-3) / 1e12
    return {"kernel": "dbev::conv3x3_halo_kernel<256>", "bound": "tensor", "achieved": round(ach, 1), "peak": round(peak, 1),
            "unit": "TFLOP/s", "frac": round(ach / peak, 4),
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the ncu --set full capture in
            # profiles/r02_conv_train.json (287.4 + 93.4 MB; the rest of the 134 MB output is still in L2 when the
            # kernel ends; algorithmic: x 268 MB + W 4.7 MB + y 134 MB)
            "traffic": 380.8e6, "peak_source": how,
            "algorithmic_flops_per_launch": flops, "launch_ms": round(ms, 5),
            "shape": "FPN_LSS conv 512 -> 256, 3x3, [8,128,128] NHWC fp32 (TF32 multiply, fp32 accumulate)",
            "ncu": "profiles/r02_conv_train.json (sm__pipe_tensor_cycles_active, per kernel)"}


def teacher_conv_roofline(hp):
    """Tensor-bound companion of `roofline`: the teacher's SECOND + SECONDFPN stack (the kernels with the
    largest share of the v2 step, conv2d_tc_kernel<64|128|256>) timed alone with CUDA events on its
    stream. peak = half of the measured bf16 rate (TF32 issues at half the bf16 rate on tcgen05)."""
    import torch
    dbev, dev = hp.dbev, hp.dev
    with torch.no_grad():
        canvas = dbev.pillar_canvas(hp.d_points, hp.enc, hp.scat)
        for _ in range(3):
            hp.secfpn(hp.second(canvas))
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        iters = 10
        a.record()
        for _ in range(iters):
            hp.secfpn(hp.second(canvas))
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    flops, hw, cin = 0.0, 512, 64
    for n, c in zip((3, 5, 5), (64, 128, 256)):
        hw //= 2
        flops += 2.0 * BATCH * hw * hw * 9 * cin * c + n * 2.0 * BATCH * hw * hw * 9 * c * c
        cin = c
    flops += 2.0 * BATCH * 128 * 128 * (4 * 64 * 128 + 128 * 128 + 256 * 128)
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak, how = float(json.load(open(path))["bf16_tflops"]) / 2.0, "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 rate)"
    else:
        peak, how = 1100.0, "fallback: nominal dense TF32 1.1 PFLOP/s"
    ach = flops / (ms * 1e-3) / 1e12
    return {"kernel": "dbev::conv3x3_halo_kernel<64|128|256> x 13 + dbev::conv2d_tc_kernel x 7 launches (SECOND + SECONDFPN forward)", "bound": "tensor",
            "achieved": round(ach, 1), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(ach / peak, 4),
            "traffic": None, "peak_source": how, "flops_per_stack": flops, "stack_ms": round(ms, 4),
            "ncu": "profiles/r01_conv3x3_halo.json (sm__pipe_tensor_cycles_active 42 / 68 / 75 % of active cycles on the 64 / 128 / 256-channel 3x3 layers)"}


def bev_pool_roofline(device):
    """Live roofline of the dominant bev_pool kernel: gather-forward over materialised frustum
    features at the configs[1] shape (16 sample-frames, C=64): kernel timed alone with CUDA events
    on the launching stream; inputs (0.9 GB) exceed L2, no flush needed."""
    import torch
    import distill_bev_b200 as dbev
    from distill_bev_b200 import _lib, synthetic
    from distill_bev_b200.plugin.ops import bev_pool as bp
    nf = BATCH * FRAMES
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=8).to(device)
    calib = [torch.from_numpy(a).to(device) for a in synthetic.make_calibration(nf, N_CAMS, seed=rank_seed(0))]
    geom = vt.get_geometry(*calib)
    plan = vt.make_plan(geom, nf, with_point_cell=False)
    n = geom.numel() // 3
    x = torch.rand(n, C_TRANS, device=device)
    lib = _lib.load()
    stream = torch.cuda.current_stream(device)

    def time_layout(layout):
        shape, sB, sZ, sC = bp._out_strides(plan, C_TRANS, layout)
        out = torch.empty(shape, device=device)

        def launch():
            rc = lib.dbev_bev_pool_gather_forward(
                _lib.ptr(x), C_TRANS, _lib.ptr(plan.order), _lib.ptr(plan.cell_start), _lib.ptr(plan.cell_end),
                _lib.ptr(plan.items), _lib.ptr(plan.n_items), plan.batch, plan.nz, plan.nslow, plan.nfast,
                sB, sZ, sC, _lib.ptr(out), _lib.stream_ptr(device))
            _lib.check(rc, "dbev_bev_pool_gather_forward")
        for _ in range(5):
            launch()
        iters = 30
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(iters):
            launch()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters, out

    ms_nchw, out_nchw = time_layout("bz_c")
    ms, out = time_layout("cl")
    same = bool(torch.equal(out.view(plan.batch, plan.nslow, plan.nfast, C_TRANS).permute(0, 3, 1, 2), out_nchw))
    kept = plan.num_kept()
    alg_bytes = kept * C_TRANS * 4 + kept * 4 + out.numel() * 4
    peak, how = measured_peaks()
    ach = alg_bytes / (ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r01_bev_pool_fwd_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    ach_nchw = alg_bytes / (ms_nchw * 1e-3) / 1e9
    return {"kernel": "dbev::bev_pool_gather_fwd_kernel<16,false,6,true> (channels-last output rows)", "bound": "hbm",
            "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic,
            "peak_source": how, "algorithmic_bytes_per_launch": int(alg_bytes),
            "launch_ms": round(ms, 5), "shape": "16 sample-frames, n=%d rows kept of %d, C=64, 128x128" % (kept, n),
            "nchw_output_variant": {"kernel": "dbev::bev_pool_gather_fwd_kernel<16,false,6,false> (transposing epilogue)",
                                    "launch_ms": round(ms_nchw, 5), "achieved": round(ach_nchw, 1),
                                    "frac": round(ach_nchw / peak, 4), "bit_identical_values": same}}


def run_configs3(args):
    """BASELINE.json configs[3] (CenterPoint/LidarFormer sparse teacher -> BEVFormer student, 200 x 200 BEV): the part of
    it that is on the SURVEY §8 hot path - the frozen sparse teacher end to end (hard voxelize 0.064 m voxels on a
    41 x 1600 x 1600 grid -> HardSimpleVFE -> SparseEncoder: 20 sparse convs, tcgen05 3xTF32 for C >= 32 ->
    [B, 256, 200, 200]) and the BEVFormer-variant distillation at the BEV position (cell-centre masks, FP mask from the
    teacher's boxes, fgd loss without channel terms, hs loss; forward + backward w.r.t. the student BEV embedding).
    The BEVFormer student's transformer itself (third-party mmcv / BEVFormer code) is not in the step. Per-GPU batch 2,
    240k-point clouds (10 sweeps). One JSON line."""
    import torch
    import distill_bev_b200 as dbev
    from distill_bev_b200 import synthetic
    from distill_bev_b200.plugin.distill import bevformer as bf
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, n_points, H = 2, 240000, 200
    vox = dbev.Voxelization([0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 10, (90000, 120000)).eval()
    vfe = dbev.HardSimpleVFE(5)
    torch.manual_seed(0)
    enc = dbev.SparseEncoder(
        in_channels=5, sparse_shape=[41, 1600, 1600], output_channels=128,
        encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
        encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type="basicblock").to(dev).eval()
    clouds = [torch.from_numpy(c).to(dev) for c in synthetic.make_lidar_scene(B, n_points, seed=rank_seed(rank))]
    gt = synthetic.make_gt_boxes(B, seed=rank_seed(rank))
    boxes = [torch.from_numpy(b) for b, _ in gt]
    g = torch.Generator().manual_seed(rank_seed(rank))
    preds = [(torch.from_numpy(b).float() + 0.3 * torch.randn(b.shape, generator=g).float(), torch.rand(len(b), generator=g), None)
             for b, _ in gt]
    student = torch.relu(torch.randn(B, 256, H, H, generator=g)).to(dev).requires_grad_(True)
    s_hs, t_hs = torch.randn(B, 256, 900, generator=g).to(dev).requires_grad_(True), torch.randn(B, 256, 900, generator=g).to(dev)
    spatial = torch.nn.Conv2d(1, 1, 3, padding=1).to(dev)
    params = dict(DISTILL_PARAMS, fp_as_foreground=["teacher"], hs_feat_loss_weights=1e-3)
    tcfg = dict(grid_size=[1600, 1600, 40], point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], voxel_size=[0.064, 0.064, 0.2])

    def step():
        with torch.no_grad():
            feats, coors = [], []
            for b, pts in enumerate(clouds):
                v, c, n = vox(pts)
                feats.append(vfe(v, n, c))
                coors.append(torch.nn.functional.pad(c, (1, 0), value=b))
            teacher = enc(torch.cat(feats), torch.cat(coors).contiguous(), B)
        losses = bf.fgd_distill_loss(teacher, student, boxes, preds, params, tcfg, spatial_adaptation=spatial, epoch=1)
        losses.update(bf.hs_distill_loss(t_hs, s_hs, params))
        sum(losses.values()).backward()
        student.grad = s_hs.grad = None
        spatial.zero_grad(set_to_none=True)
        return losses

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(args.steps):
        losses = step()
    b.record()
    torch.cuda.synchronize()
    ms = aggregate_step_time(a.elapsed_time(b), world, dist if world > 1 else None, dev) / args.steps
    if rank != 0:
        return
    line = {"metric": METRIC, "value": round(B * world / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (sparse convs 3xTF32 on tcgen05: fp32-equivalent)", "data": "synthetic",
            "config": {"workload": "configs[3] hot-path part: frozen sparse LiDAR teacher (2 x 240k pts -> hard voxelize -> HardSimpleVFE "
                                   "-> LidarFormer SparseEncoder -> [2,256,200,200]) + BEVFormer-variant fgd / hs distillation loss "
                                   "fwd+bwd at 200x200x256; eager (the voxel / rulebook counts are read back as the reference API implies); "
                                   "BEVFormer student transformer not in the step", "batch_per_gpu": B, "parallelism": "dp%d" % world},
            "losses_finite": bool(all(torch.isfinite(v).item() for v in losses.values())),
            "sparse_teacher": sparse_teacher_probe(dev, batch=B, n_points=n_points),
            "roofline": sparse_conv_roofline(dev, batch=B, n_points=n_points)}
    print(json.dumps(line))


def run_configs4(args):
    """BASELINE.json configs[4]: the bev_pool + distill-loss sweep (tools/sweep_configs4.py) as one JSON line; value = the
    channels-last gather's GB/s at the configs[1]-like point (D=59, BEV 128, C=64... reported per point in `sweep`)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_configs4
    sweep_configs4.main()
    res = json.load(open(os.path.join(ROOT, "gpurun_out", "configs4_sweep.json")))
    best = max(r["gather_channels_last_hbm_frac"] for r in res["bev_pool"])
    line = {"metric": "bev_pool_gather_hbm_fraction_sweep", "value": best, "unit": "fraction of measured HBM peak (best point)",
            "n_gpus": 1, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4] sweep: D in {59,118} x BEV in {128,256,512} x C in {64,256}; 4 sample-frames per launch"},
            "sweep": res}
    print(json.dumps(line))


def _exit_multi_rank():
    """Leave without tearing the NCCL communicator down: CUDA graphs that recorded NCCL kernels keep it busy and
    destroy_process_group() can wait forever on them. All results are printed and flushed before this is called."""
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    hp = HotPath(device, seed=rank_seed(rank), world=world, allreduce=args.allreduce, comm_dtype=args.comm_dtype)
    hp.sorted_splat = args.sorted_splat
    graph_note = "eager (--no-graph)"
    if not args.no_graph:
        attempts = [args.allreduce] + (["after"] if (world > 1 and args.allreduce == "overlap") else [])
        for mode in attempts:
            try:
                if mode != hp.allreduce:
                    hp.set_allreduce(mode)
                hp.enable_graph()
                graph_note = ("step captured once in a CUDA graph and replayed%s; e2e copies every batch from pinned host memory "
                              "into one of two static input sets on a copy stream while the previous step runs (prefetch), "
                              "and reads the losses back every step"
                              % (" (forward+backward graph, eager NCCL all-reduce, optimizer graph)" if hp.split else
                                 (" (NCCL all-reduce kernels recorded in the graph, launched per bucket from gradient hooks)"
                                  if hp.reducer is not None else "")))
                break
            except Exception as exc:  # keep measuring, eagerly, and say so
                hp.captured = None
                graph_note = "eager (graph capture failed: %s)" % str(exc).splitlines()[0][:120]
                sys.stderr.write("bench.py: CUDA graph capture failed (allreduce=%s): %s\n" % (mode, exc))
                import traceback
                traceback.print_exc(file=sys.stderr)
                torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e):
        for _ in range(max(args.warmup, 3)):
            hp.step(e2e)
        barrier()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(args.steps):
            hp.step(e2e)
        b.record()
        torch.cuda.synchronize()
        ms = aggregate_step_time(a.elapsed_time(b), world, dist, device)
        barrier()
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = timed(False)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = timed(True)
    ours, all_k = count_our_kernels(hp)      # on every rank: the eager step contains the gradient all-reduce
    if rank != 0:
        if world > 1:
            dist.barrier()
            _exit_multi_rank()
        return
    if world > 1:
        dist.barrier()
    ms_per_step = total_ms / args.steps
    value = samples_per_sec(total_ms, args.steps, world)
    e2e_value = samples_per_sec(e2e_ms, args.steps, world)
    flops = 3.0 * encoder_flops(BATCH)
    coll = hp.reducer.describe() if hp.reducer is not None else {"collective": "none (N = 1)" if world == 1 else "none (--allreduce none)"}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (TF32 tensor-core multiply, fp32 accumulate: the reference's "
                                                         "cuDNN arithmetic under torch defaults)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "frames": FRAMES, "parallelism": "dp%d" % world,
                   "l2": "inputs larger than L2 (activations > 3 GB per step)",
                   "collective": "%s, %d bytes per step (%s)" % (coll["collective"], coll.get("bytes_per_step", 0),
                                                                 coll.get("comm_dtype", "-"))},
        "step_detail": {"l2": "activations larger than L2 (encoder activations > 3 GB, teacher conv activations > 2 GB per step)",
                        "collective": coll, "allreduce_mode": hp.allreduce if world > 1 else "n/a",
                        "trainable_parameters": int(hp.n_params), "optimizer": "torch.optim.AdamW(fused=True, capturable=True)",
                        "student_encoder_train_flops_per_step": flops,
                        "student_encoder_tflops_if_alone": None,
                        "cuda_graph": hp.captured is not None, "issue": graph_note},
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(hp.h2d_bytes),
                "d2h_bytes_per_step": int(hp.d2h_bytes),
                "note": "host-originated inputs (calibration, LiDAR, GT boxes + labels) copied from "
                        "pinned memory every step, loss scalars read back"},
        "gpu_launches": int(ours * args.steps), "gpu_launches_per_step": int(ours),
        "all_cuda_kernels_per_step": all_k, "clocks": clocks,
    }
    del line["step_detail"]["student_encoder_tflops_if_alone"]
    if world == 1:
        try:
            line["roofline"] = student_conv_roofline(hp)
        except Exception as exc:  # noqa: BLE001
            line["roofline"] = {"error": str(exc)[:200]}
        line["roofline_bev_pool"] = bev_pool_roofline(device)
        try:
            line["roofline_teacher_convs"] = teacher_conv_roofline(hp)
        except Exception as exc:
            line["roofline_teacher_convs"] = {"error": str(exc)[:200]}
        line["cpu_baseline"] = cpu_baseline(samples=1, procs=1)
        for key, probe in (("cudnn_encoder_step", lambda: cudnn_encoder_probe(device, rank_seed(rank))),
                           ("sparse_teacher", lambda: sparse_teacher_probe(device))):
            try:
                line[key] = probe()
            except Exception as exc:  # extra evidence only: never lose the headline line over it
                line[key] = {"error": str(exc)[:200]}
    else:
        for key, probe in (("roofline", lambda: student_conv_roofline(hp)), ("roofline_bev_pool", lambda: bev_pool_roofline(device))):
            try:
                line[key] = probe()
            except Exception as exc:  # never lose the headline line over a probe
                line[key] = {"error": str(exc)[:200]}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        _exit_multi_rank()


# ----------------------------------------------------------------------------- CPU oracle arm

def _cpu_one_sample(seed):
    """The same three stages for ONE sample on the CPU oracle port (numpy + C)."""
    synthetic = _load_synthetic()
    from oracle import fgd_oracle, lss_oracle, pillar_oracle, voxel_oracle
    rng = np.random.RandomState(seed)
    t0 = time.perf_counter()
    # A: 2 frames of one sample
    grid = synthetic.NUSC_GRID
    dx, bx, nx = lss_oracle.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    frustum = lss_oracle.create_frustum(synthetic.NUSC_INPUT_SIZE, 16, grid["dbound"])
    calib = synthetic.make_calibration(FRAMES, N_CAMS, seed=seed)
    geom = lss_oracle.get_geometry(frustum, *calib)
    dl = rng.randn(FRAMES * N_CAMS, D, FH, FW).astype(np.float32)
    depth = np.exp(dl) / np.exp(dl).sum(1, keepdims=True)
    feat = rng.randn(FRAMES * N_CAMS, C_TRANS, FH, FW).astype(np.float32)
    bev = lss_oracle.lift_splat(geom, depth, feat, FRAMES, N_CAMS, bx, dx, nx)
    lss_oracle.lift_splat_backward(geom, depth, feat, np.ones_like(bev), FRAMES, N_CAMS, bx, dx, nx)
    # B: one cloud
    pts = synthetic.make_lidar(1, N_POINTS, seed=seed)[0]
    coors = voxel_oracle.dynamic_voxelize(pts, PILLAR_VS, PILLAR_RANGE)
    coors = np.concatenate([np.zeros((pts.shape[0], 1), np.int32), coors], 1)
    w = rng.randn(64, 10).astype(np.float32) * 0.1
    vf, vc = pillar_oracle.pillar_encode(pts, coors, w, np.ones(64), np.zeros(64), np.zeros(64), np.ones(64),
                                         1e-3, PILLAR_VS, PILLAR_RANGE)
    canvas = pillar_oracle.pillar_scatter(vf, vc, 1, 512, 512)
    _cpu_teacher_convs(canvas)
    # S: student BEV encoder, training mode, forward + backward (torch.nn on the CPU = the reference's implementation)
    _cpu_student_encoder(np.concatenate([bev[0:1], bev[1:2]], 1))
    # C: one sample at the head position
    teacher = np.maximum(rng.randn(1, C_TEACHER, BEV, BEV), 0).astype(np.float32)
    student = np.maximum(rng.randn(1, C_STUDENT, BEV, BEV), 0).astype(np.float32)
    wa = (rng.randn(C_TEACHER, C_STUDENT) * 0.05).astype(np.float32)
    adapted = np.einsum("oc,bchw->bohw", wa, student, optimize=True).astype(np.float32)
    boxes = [synthetic.make_gt_boxes(1, seed=seed)[0][0]]
    fg, fgs, bgs = fgd_oracle.foreground_scale_mask(BEV, BEV, boxes, TRAIN_CFG["grid_size"],
                                                    TRAIN_CFG["point_cloud_range"], TRAIN_CFG["voxel_size"])
    gt_hm = (rng.random_sample((1, 10, BEV, BEV)) ** 12).astype(np.float32)
    tl = (rng.randn(1, 10, BEV, BEV) * 1.5 - 3.0)
    sig = np.clip(1 / (1 + np.exp(-tl)), 1e-4, 1 - 1e-4).astype(np.float32)
    fp, fps, cnt = fgd_oracle.add_fp_as_fg("teacher", fg, gt_hm, sig, np.zeros_like(sig), 0.1)
    p = dict(spatial_t=0.5, spatial_student_ratio=1.0, channel_t=0.5, w_fg=6e-3, w_bg=4e-2, w_channel=0.25,
             w_spatial=2.5e-3, w_fp=6e-2, spatial_att="teacher_student", spatial_mask=True, channel_mask=False,
             scale_mask="combine_gt")
    res = fgd_oracle.fgd_loss(teacher, adapted, fg, fgs, bgs, p, conv_w=np.full((3, 3), 0.1), conv_b=0.0,
                              fp=fp, fp_scale=fps, fp_count=cnt, want_grad=True)
    np.einsum("oc,bohw->bchw", wa, res["grad_student"].astype(np.float32), optimize=True)
    return time.perf_counter() - t0


_CPU_TEACHER = None


def _cpu_teacher_convs(canvas):
    """SECOND + SECONDFPN forward of one sample on the host: the reference's own implementation of
    these rows is torch.nn (Conv2d / BatchNorm2d / ReLU / ConvTranspose2d built by mmcv's
    build_conv_layer), so the CPU arm runs exactly that, one thread per worker process."""
    global _CPU_TEACHER
    import torch
    torch.set_num_threads(int(os.environ.get("DBEV_CPU_THREADS", "1")))
    if _CPU_TEACHER is None:
        nn = torch.nn
        torch.manual_seed(0)
        blocks, cin = [], 64
        for n, c in zip((3, 5, 5), (64, 128, 256)):
            layers = [nn.Conv2d(cin, c, 3, 2, 1, bias=False), nn.BatchNorm2d(c, eps=1e-3), nn.ReLU(inplace=True)]
            for _ in range(n):
                layers += [nn.Conv2d(c, c, 3, 1, 1, bias=False), nn.BatchNorm2d(c, eps=1e-3), nn.ReLU(inplace=True)]
            blocks.append(nn.Sequential(*layers))
            cin = c
        de = [nn.Sequential(nn.Conv2d(64, 128, 2, 2, bias=False), nn.BatchNorm2d(128, eps=1e-3), nn.ReLU(inplace=True)),
              nn.Sequential(nn.ConvTranspose2d(128, 128, 1, 1, bias=False), nn.BatchNorm2d(128, eps=1e-3), nn.ReLU(inplace=True)),
              nn.Sequential(nn.ConvTranspose2d(256, 128, 2, 2, bias=False), nn.BatchNorm2d(128, eps=1e-3), nn.ReLU(inplace=True))]
        _CPU_TEACHER = (nn.ModuleList(blocks).eval(), nn.ModuleList(de).eval())
    blocks, de = _CPU_TEACHER
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(canvas, dtype=np.float32)).reshape(1, 64, 512, 512)
        outs = []
        for b in blocks:
            x = b(x)
            outs.append(x)
        return torch.cat([d(o) for d, o in zip(de, outs)], 1)


_CPU_STUDENT = None


def _cpu_student_encoder(bev):
    """ResNetForBEVDet + FPN_LSS forward and backward of one sample on the host (nn.Conv2d / BatchNorm2d in
    training mode / ReLU / Upsample - the modules the reference builds), one thread per worker process."""
    global _CPU_STUDENT
    import torch
    torch.set_num_threads(int(os.environ.get("DBEV_CPU_THREADS", "1")))
    if _CPU_STUDENT is None:
        _CPU_STUDENT = CudnnEncoder.build(torch.device("cpu")).to(memory_format=torch.contiguous_format)
    x = torch.from_numpy(np.ascontiguousarray(bev, dtype=np.float32)).reshape(1, 2 * C_TRANS, BEV, BEV).requires_grad_(True)
    y = _CPU_STUDENT(x)
    y.backward(torch.ones_like(y))
    _CPU_STUDENT.zero_grad(set_to_none=True)
    return y.detach()


def _load_synthetic():
    """distill-bev_b200/synthetic.py as a standalone module (numpy only; keeps torch out of the
    CPU worker processes)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_dbev_synthetic", os.path.join(ROOT, "distill-bev_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_baseline(samples, procs, pool=None):
    """samples/s of the CPU oracle port on `procs` host processes (bounded sample)."""
    t0 = time.perf_counter()
    if pool is None:
        for i in range(samples):
            _cpu_one_sample(7 + i)
    else:
        pool.map(_cpu_one_sample, [7 + i for i in range(samples)])
    dt = time.perf_counter() - t0
    return {"value": round(samples / dt, 4), "unit": UNIT, "cores": procs, "kind": "port",
            "sample": "%d sample(s): the same stages (2 frames lift+splat fwd/bwd, student BEV encoder fwd/bwd and one "
                      "30k-point cloud through pillar encoder + SECOND/SECONDFPN [torch.nn on the CPU, 1 thread per process], "
                      "one head-position loss fwd/bwd incl. the 1x1 adaptation) on oracle/ (numpy + C) + torch.nn; "
                      "host has %d cores" % (samples, os.cpu_count() or 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import multiprocessing as mp
    # one sample per process, 4 threads per process for the torch.nn conv stacks (no oversubscription): a step stays
    # a few seconds long, so the driver's --steps 20 --warmup 5 run ends within a few minutes
    cores = max(1, min(os.cpu_count() or 1, 16))
    threads = 4 if cores >= 4 else 1
    procs = max(1, cores // threads)
    per_step = procs
    os.environ["DBEV_CPU_THREADS"] = str(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", str(threads))
    with mp.get_context("spawn").Pool(procs) as pool:
        for _ in range(max(args.warmup, 0)):
            cpu_baseline(per_step, procs, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = cpu_baseline(per_step, procs, pool)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    res["value"] = round(value, 4)
    res["cores"] = procs * threads
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 0), "ms_per_step": round(dt / args.steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "frames": FRAMES, "parallelism": "dp%d" % world},
            "step_detail": {"note": "CPU oracle port of the reference algorithm (the reference's Python files need "
                                    "/root/reference + mmcv and cannot travel to the GPU box) + torch.nn conv stacks on the CPU; "
                                    "each step = %d samples in %d host processes; no optimizer step" % (per_step, procs)},
            "cpu_baseline": res,
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly (no CUDA graph replay)")
    ap.add_argument("--allreduce", default="overlap", choices=["after", "overlap", "none"],
                    help="N > 1: gradient all-reduce recorded into the step's CUDA graph and overlapped with the backward "
                         "(default; falls back to 'after' if the capture fails), issued eagerly after the backward graph "
                         "('after'), or skipped ('none': replicas, not training)")
    ap.add_argument("--comm-dtype", default="bf16", choices=["bf16", "f32"], help="wire dtype of the gradient all-reduce")
    ap.add_argument("--sorted-splat", action="store_true",
                    help="lift+splat through the sorted plan (fixed summation order) instead of the sort-free splat")
    ap.add_argument("--config", default="configs1", choices=["configs1", "configs3", "configs4"],
                    help="BASELINE.json config: configs1 = the headline training step (default, what the driver runs); "
                         "configs3 = sparse teacher + BEVFormer-variant loss at 200 x 200; configs4 = the bev_pool / loss sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "configs3":
        run_configs3(args)
    elif args.config == "configs4":
        run_configs4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
