"""Data-parallel training runtime: one process per GPU, replicated weights, ONE NCCL all-reduce per step.

Replaces the reference's multi-GPU machinery: torch.nn.DataParallel's per-iteration
parameter broadcast / logits gather / gradient reduce-add onto GPU 0 (zs3/train_pascal.py:90-93) and the
per-layer SyncBN reduce+broadcast pairs (zs3/modeling/sync_batchnorm/batchnorm.py:101-122).  Here:

  * all parameters are views into one flat fp32 buffer, all gradients views into a second one;
  * after backward a single ncclAllReduce(sum) over the flat gradient buffer runs over NVLink/NVSwitch;
  * a fused SGD kernel (csrc/misc.cu) updates the flat parameter buffer, folding in the 1/world_size.

BatchNorm statistics stay rank-local (16 images per GPU), a documented deviation from the reference's
SyncBN-on-multi-GPU choice (DESIGN.md "Multi-GPU").
"""
import os

import torch
import torch.distributed as dist

from . import functional as ZF
from . import kernels as K


def init_distributed():
    """Initialise torch.distributed from torchrun's environment; returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # test knobs: ZS3_DIST_BACKEND=gloo and ZS3_DEVICE_INDEX=0 let several ranks share one GPU (NCCL refuses that),
    # which exercises the multi-rank control flow on a single-GPU box
    if "ZS3_DEVICE_INDEX" in os.environ:
        local_rank = int(os.environ["ZS3_DEVICE_INDEX"])
    if world > 1 and not dist.is_initialized():
        backend = os.environ.get("ZS3_DIST_BACKEND", "nccl" if torch.cuda.is_available() else "gloo")
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend)
    return rank, local_rank, world


def shard_batch(n_total, rank, world):
    """contiguous [begin, end) slice of a global batch owned by `rank` (images shard, weights replicate)"""
    per = n_total // world
    rem = n_total % world
    begin = rank * per + min(rank, rem)
    return begin, begin + per + (1 if rank < rem else 0)


def exchange_step2(generator_params, snapshot, classifier_params, world):
    """The ONE collective of a data-parallel ZS3Net step-2 iteration (the reference runs step 2 on a single
    GPU, train_pascal_GMMN.py:155,262, so these semantics are this repository's, stated in DESIGN.md "Multi-GPU"):
    every rank has run its own sequential generator chain on its own images from the same starting weights
    (`snapshot`, one flat tensor) and the classifier backward on its own batch.  One all-reduce(sum) over
    [generator delta | classifier gradients] (220 k + 5.4 k floats at 21 classes: < 1 MB); afterwards every rank holds
    generator = snapshot + mean delta and classifier .grad = mean gradient.  Adam moments stay rank-local.
    Returns the flat message (for tests)."""
    with torch.no_grad():
        cur = torch.cat([p.detach().reshape(-1) for p in generator_params])
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in classifier_params]
        msg = torch.cat([cur - snapshot] + [g.reshape(-1).to(cur.dtype) for g in grads])
        if world > 1:
            dist.all_reduce(msg)
            msg /= world
        off = 0
        for p in generator_params:
            n = p.numel()
            p.copy_((snapshot[off:off + n] + msg[off:off + n]).view_as(p))
            off += n
        for p, g in zip(classifier_params, grads):
            n = p.numel()
            g.copy_(msg[off:off + n].view_as(g))
            p.grad = g
            off += n
    return msg


class FlatParams:
    """Re-homes the parameters of `param_groups` (list of lists) into one flat buffer (+ a flat grad buffer)."""

    def __init__(self, param_groups):
        params = [p for g in param_groups for p in g]
        if not params:
            raise ValueError("no parameters")
        dev, dt = params[0].device, params[0].dtype
        total = sum(p.numel() for p in params)
        self.flat = torch.empty(total, dtype=dt, device=dev)
        self.grad = torch.zeros(total, dtype=dt, device=dev)
        self.group_ranges = []
        off = 0
        for g in param_groups:
            start = off
            for p in g:
                n = p.numel()
                # keep each parameter's memory layout (conv weights are KRSC / channels_last): dense strided views
                view = torch.as_strided(self.flat, p.shape, p.stride(), off)
                view.copy_(p.detach())
                p.data = view
                p.grad = torch.as_strided(self.grad, p.shape, p.stride(), off)
                off += n
            self.group_ranges.append((start, off))
        self.params = params

    def zero_grad(self):
        self.grad.zero_()
        # autograd accumulates in place into the existing .grad views
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr():
                raise RuntimeError("parameter gradient was detached from the flat buffer (use zero_grad of this runtime)")


class FusedSGD:
    """torch.optim.SGD semantics (momentum, weight decay, nesterov, per-group lr) over FlatParams ranges."""

    def __init__(self, flat, lrs, momentum=0.9, weight_decay=5e-4, nesterov=False):
        self.flat = flat
        self.momentum, self.weight_decay, self.nesterov = momentum, weight_decay, nesterov
        self.buf = torch.zeros_like(flat.flat)
        self.steps = 0
        # per-group learning rates live in device memory and are read by the kernel when it RUNS, so a step captured in
        # a CUDA graph keeps following set_lr() (the reference applies its poly schedule every iteration,
        # base_trainer.py:15 -> lr_scheduler.py:46-67: group 0 = lr, groups > 0 = 10 * lr)
        self._lrs = [float(v) for v in lrs]
        self.lr_dev = torch.tensor(self._lrs, dtype=torch.float32, device=flat.flat.device)

    @property
    def lrs(self):
        return list(self._lrs)

    @lrs.setter
    def lrs(self, values):
        self.set_lr(values)

    def set_lr(self, values):
        """per-group learning rates (a list, or one float = the schedule's lr: group 0 gets lr, the others 10 * lr as
        LR_Scheduler.__call__ does).  Takes effect on the next step, also under CUDA-graph replay."""
        if not isinstance(values, (list, tuple)):
            values = [float(values)] + [float(values) * 10.0] * (len(self._lrs) - 1)
        if len(values) != len(self._lrs):
            raise ValueError(f"{len(self._lrs)} parameter groups, {len(values)} learning rates")
        self._lrs = [float(v) for v in values]
        self.lr_dev.copy_(torch.tensor(self._lrs, dtype=torch.float32), non_blocking=True)

    def step(self, grad_scale=1.0):
        self.step_span(0, self.flat.flat.numel(), grad_scale)
        self.finish_step()

    def step_span(self, lo, hi, grad_scale=1.0):
        """the update of the flat elements [lo, hi) only (each parameter group's part with its own learning rate): lets
        the data-parallel runtime update the parameters whose gradients are already reduced while the rest of the
        backward still runs; finish_step() closes the iteration once every span has been issued"""
        for gi, (a, b) in enumerate(self.flat.group_ranges):
            a, b = max(a, lo), min(b, hi)
            if b > a:
                K.sgd_step(self.flat.flat[a:b], self.flat.grad[a:b], self.buf[a:b], self.lr_dev[gi:gi + 1], self.momentum,
                           self.weight_decay, self.nesterov, False, grad_scale)  # buf starts at 0: mom*0 + d == torch's first-step buf = d

    def finish_step(self):
        self.steps += 1
        ZF.invalidate_weight_caches()


class HostPrefetcher:
    """Double-buffered host->device staging of (image, target) batches on a side stream, so that the PCIe copy of
    batch i+1 overlaps the training step of batch i (the reference gets this from DataLoader(pin_memory=True) +
    .cuda(); zs3/base_trainer.py:12-14).  Usage: stage(0, batch0); for i: a, b = take(i); stage(i+1, next); step(a, b)."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [None, None]

    def stage(self, i, host_batch):
        k = i & 1
        with torch.cuda.stream(self.stream):
            if self.consumed[k] is not None:
                self.stream.wait_event(self.consumed[k])  # the step that read this slot has finished reading it
            if self.slots[k] is None:
                self.slots[k] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_batch)
            for dst, src in zip(self.slots[k], host_batch):
                dst.copy_(src, non_blocking=True)
            self.ready[k].record(self.stream)

    def take(self, i):
        k = i & 1
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        return self.slots[k]

    def release(self, i):
        """call after the consumer has enqueued its reads of slot i (e.g. after train_step)"""
        k = i & 1
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.consumed[k] = ev


def cut_backward(loss, cut, early_params):
    """Two-stage backward.  Stage 1 (run here): gradients of `loss` w.r.t. the `cut` activations and the parameters
    above them (`early_params`; accumulated into .grad -- by the fused nodes themselves where they write the weight
    gradient in place, otherwise here).  Returns tail(), which backpropagates the cut gradients through everything
    below.  Two requirements, both learnt from the first 2-GPU run of this path:
      * the cut must be an antichain of the autograd graph (no cut tensor upstream of another);
      * stage 1 must CAPTURE the gradients at the cut (torch.autograd.grad): `backward(inputs=[non-leaf])` executes
        the node that produced the tensor, which the tail then finds already run (and, for the fused nodes, freed)."""
    early_params = list(early_params)
    got = torch.autograd.grad(loss, list(cut) + early_params, allow_unused=True)
    grads = list(got[:len(cut)])
    with torch.no_grad():
        for p, g in zip(early_params, got[len(cut):]):
            if g is not None:  # (None: the node accumulated into p.grad itself)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.add_(g)

    def tail():
        live = [(t, g) for t, g in zip(cut, grads) if g is not None]
        torch.autograd.backward([t for t, _ in live], [g for _, g in live])

    return tail


class DataParallelTrainer:
    """step-1 training step of zs3/base_trainer.py:16-20 (zero_grad, forward, CE, backward, SGD) for one rank."""

    def __init__(self, model, criterion, lr=0.007, momentum=0.9, weight_decay=5e-4, nesterov=False, world_size=1,
                 use_cuda_graph=False, fuse_loss=True, sync_bn=False):
        """sync_bn: synchronise the training-mode statistics of every SynchronizedBatchNorm2d over the ranks (the
        reference enables its SyncBN whenever it trains on more than one GPU, train_pascal.py:279); default False =
        rank-local statistics.  The per-layer statistic all-reduces sit between kernels, so sync_bn runs eagerly."""
        self.model, self.criterion, self.world = model, criterion, world_size
        self.fuse_loss = fuse_loss
        self.sync_bn = bool(sync_bn) and world_size > 1
        from .modeling.sync_batchnorm.batchnorm import enable_sync
        enable_sync(world_size if self.sync_bn else 1)
        if self.sync_bn and os.environ.get("ZS3_SYNCBN_GRAPH", "0") != "1":
            use_cuda_graph = False
        # Loss scaling.  Every rank's criterion divides by ITS batch size and weight sum (loss.py:43-44), the reference's
        # DataParallel evaluates ONE criterion on the gathered global batch (train_pascal.py:90-93 + base_trainer.py:18):
        # sum_r grad(L_r) = world^2 * grad(L_global) for equally sized shards, hence 1/world^2 on the summed gradient.
        self.grad_scale = 1.0 / float(world_size * world_size)
        self.use_cuda_graph, self.graph, self.graph_tail = use_cuda_graph, None, None
        groups = [list(model.get_1x_lr_params()), list(model.get_10x_lr_params())]
        self.flat = FlatParams(groups)
        self.opt = FusedSGD(self.flat, [lr, lr * 10], momentum, weight_decay, nesterov)
        ZF.invalidate_weight_caches()
        # bf16 shadow of the flat parameter buffer: conv weights are KRSC, so a layer's slice of the shadow is its
        # packed forward weight [cout][R*S][cin] (layers with channel padding keep their own packed copies)
        self.shadow = torch.empty(self.flat.flat.numel(), dtype=torch.bfloat16, device=self.flat.flat.device)
        K.cast_f32_to_bf16(self.flat.flat, self.shadow)   # valid from construction on; _begin_step refreshes it per step
        base = self.flat.flat.data_ptr()
        for m in model.modules():
            if isinstance(m, torch.nn.Conv2d) and K.is_krsc(m.weight):
                off = (m.weight.data_ptr() - base) // 4
                if 0 <= off < self.flat.flat.numel() and off % 8 == 0:
                    co, ci, r, s = m.weight.shape
                    m.__dict__["_zs3_bf16_shadow"] = (self.shadow[off:off + m.weight.numel()].view(co, r * s, ci),
                                                      m.weight.data_ptr())
        # parameters above the backbone cut (layer3 .. decoder): one contiguous range of the flat buffers, reduced
        # while the rest of the backward still runs (see _finish_distributed)
        # Status: ON (ZS3_DP_CUT=0 restores one all-reduce after the whole backward).  The round-1 version failed its
        # first 2-GPU run (profiles/r01_dp_cut_failure.md); cut_backward() now captures the cut gradients with
        # torch.autograd.grad and the backbone hands the decoder an alias of low_level_feat.  Validated on 2x B200
        # (tests/test_multigpu_gpu.py: gradients equal the plain flow to 2e-8; profiles/r02_dp_cut.md: A/B bench).
        self.early_range, self.early_params, self._opt_stream = None, [], None
        self.segment_events = None   # set to [] to collect (start, head done, tail done, step done) events per step
        bb = getattr(model, "backbone", None)
        if world_size > 1 and bb is not None and hasattr(bb, "layer3") and os.environ.get("ZS3_DP_CUT", "1") == "1":
            first = next(iter(bb.layer3.parameters()), None)
            if first is not None and first.grad is not None:
                a = (first.grad.data_ptr() - self.flat.grad.data_ptr()) // self.flat.grad.element_size()
                below = {id(p) for m in (bb.conv1, bb.bn1, bb.layer1, bb.layer2) for p in m.parameters()}
                self.early_params = [p for p in self.flat.params if id(p) not in below]
                early_n = sum(p.numel() for p in self.early_params)
                if 0 < a and a + early_n == self.flat.grad.numel():  # layout is [stem, layer1, layer2 | the rest]
                    self.early_range = (a, self.flat.grad.numel())
                    bb.expose_cut = True   # the backbone publishes the cut activations (backbone/resnet.py: last_cut)
        if world_size > 1:
            # identical replicas: broadcast rank 0's weights and BN buffers once
            dist.broadcast(self.flat.flat, 0)
            for b in model.buffers():
                dist.broadcast(b, 0)
            K.cast_f32_to_bf16(self.flat.flat, self.shadow)

    def train_step(self, image, target):
        """One optimisation step; returns the (device) loss tensor.  With use_cuda_graph the step is captured once
        and replayed: on one GPU a single graph (forward, loss, backward, optimizer); with several ranks two graphs,
        split inside the backward pass, with the NCCL all-reduces and the optimizer issued eagerly between/after
        them (keeps NCCL out of stream capture and overlaps the big all-reduce with the tail of the backward)."""
        if not self.use_cuda_graph:
            return self._step(image, target)
        if self.graph is None:
            self._capture(image, target)
        self.static_image.copy_(image, non_blocking=True)
        self.static_target.copy_(target, non_blocking=True)
        marks = self.segment_events
        if marks is not None:      # debug aid (tools/scale_probe.py): CUDA events around the segments of a step
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        self.graph.replay()
        if marks is not None:
            ev[1].record()
        if self.world > 1:
            tail = self.graph_tail.replay if self.graph_tail is not None else None
            if marks is not None and tail is not None:
                def tail():   # noqa: F811
                    self.graph_tail.replay()
                    ev[2].record()
            self._finish_distributed(tail)
        if marks is not None:
            if self.world == 1 or self.graph_tail is None:
                ev[2].record()
            ev[3].record()
            marks.append(ev)
        return self.static_loss

    def _capture(self, image, target):
        dev = image.device
        self.static_image, self.static_target = image.clone(), target.clone()
        ZF._RngState.device_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # the warm-up steps below are real optimisation steps on the first batch; the training state (parameters,
        # momentum, BatchNorm running statistics, step counter) is put back afterwards so that capture is invisible
        snap = (self.flat.flat.clone(), self.opt.buf.clone(), [b.clone() for b in self.model.buffers()], self.opt.steps)
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside capture: lazy inits (cudaFuncSetAttribute, scratch buffers, NCCL)
                self._step(self.static_image, self.static_target)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            self.flat.flat.copy_(snap[0])
            self.opt.buf.copy_(snap[1])
            for b, s0 in zip(self.model.buffers(), snap[2]):
                b.copy_(s0)
        self.opt.steps = snap[3]
        ZF.invalidate_weight_caches()
        del snap
        self.graph, self.graph_tail = torch.cuda.CUDAGraph(), None
        if self.world == 1:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._step(self.static_image, self.static_target)
            return
        with torch.cuda.graph(self.graph):
            self.static_loss, tail = self._forward_backward_head(self.static_image, self.static_target)
        if tail is not None:
            self.graph_tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_tail, pool=self.graph.pool()):
                tail()

    # ------------------------------------------------------------------------------------------ one step
    def _begin_step(self):
        if ZF._RngState.device_counter is not None:
            ZF._RngState.device_counter.add_(1 << 32)  # fresh dropout masks per step, also under graph replay
        self.flat.zero_grad()
        K.cast_f32_to_bf16(self.flat.flat, self.shadow)  # one launch refreshes every layer's bf16 forward weight

    def _step(self, image, target):
        if self.world > 1:
            loss, tail = self._forward_backward_head(image, target)
            self._finish_distributed(tail)
            return loss
        self._begin_step()
        loss = self._forward_loss(image, target)
        loss.backward()
        self.opt.step()
        return loss

    def _forward_backward_head(self, image, target):
        """Forward, loss and the backward of everything ABOVE the backbone's cut (layer3, layer4, ASPP, decoder).
        Returns (loss, tail) where tail() runs the rest of the backward (layer2, layer1, stem) -- or None when the
        model exposes no cut, in which case the whole backward has run."""
        self._begin_step()
        loss = self._forward_loss(image, target)
        cut = getattr(getattr(self.model, "backbone", None), "last_cut", None)
        if self.early_range is None or cut is None or not all(t.requires_grad for t in cut):
            loss.backward()
            return loss, None
        return loss, cut_backward(loss, list(cut), self.early_params)

    def _finish_distributed(self, tail):
        """all-reduce (sum) + optimizer.  With a cut: the gradients above it (one contiguous range of the flat
        buffer) are reduced on NCCL's stream WHILE the tail of the backward runs; the small remainder follows."""
        if tail is None:
            dist.all_reduce(self.flat.grad)
            self.opt.step(grad_scale=self.grad_scale)
            return
        a, b = self.early_range
        main = torch.cuda.current_stream()
        early = dist.all_reduce(self.flat.grad[a:b], async_op=True)
        tail()
        # the parameters above the cut (97 % of them) are updated on a side stream as soon as their all-reduce has
        # landed, concurrently with the tail of the backward, which only reads the weights BELOW the cut
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream()
        with torch.cuda.stream(self._opt_stream):
            early.wait()
            self.opt.step_span(a, b, self.grad_scale)
        for lo, hi in ((0, a), (b, self.flat.grad.numel())):
            if hi > lo:
                dist.all_reduce(self.flat.grad[lo:hi])
                self.opt.step_span(lo, hi, self.grad_scale)
        main.wait_stream(self._opt_stream)
        self.opt.finish_step()

    def _forward_loss(self, image, target):
        """criterion(model(image), target) (base_trainer.py:17-18).  When the criterion is this package's
        CrossEntropyLoss the x4 upsample of the class scores is fused into the loss kernels (identical value, the
        354 MB logits tensor and its gradient are never materialised)."""
        owner = getattr(self.criterion, "__self__", None)
        fusable = (self.fuse_loss and hasattr(self.model, "forward_scores")
                   and getattr(self.criterion, "__func__", None) is getattr(type(owner), "CrossEntropyLoss", None)
                   and hasattr(owner, "UpsampledCrossEntropyLoss") and getattr(self.model, "num_classes", 99) <= 64
                   and target.shape[-1] <= 640)  # limits of zs3_upsample_ce_* (classes in registers, a column per thread)
        if fusable:
            return owner.UpsampledCrossEntropyLoss(self.model.forward_scores(image), self.model.num_classes, target)
        return self.criterion(self.model(image), target)
