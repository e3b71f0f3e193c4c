"""ZS3Net step-2 training iteration (generator + classifier) on the B200 modules.

Follows the body of Trainer.training in zs3/train_pascal_GMMN.py:152-268 (identical in train_context_GMMN.py):
frozen-backbone feature extraction, per-(image, class) generator updates on 128 sampled pixels with the MMD
loss, fake/real feature assembly, classifier (pred_conv) update.  The reference trainer can drive the same
modules unchanged; this class is the in-repo runner used by tests and benchmarks.
"""
import numpy as np
import torch
from torch import nn


class ZS3Step:
    def __init__(self, model, generator, criterion, criterion_generator, optimizer, optimizer_generator,
                 seen_classes, unseen_classes, noise_dim=300, embed_dim=300, feature_dim=256,
                 batch_size_generator=128, real_seen_features=True, noise_fn=None, index_fn=None, mask_fn=None):
        self.model, self.generator = model, generator
        self.criterion, self.criterion_generator = criterion, criterion_generator
        self.optimizer, self.optimizer_generator = optimizer, optimizer_generator
        self.seen, self.unseen = set(int(c) for c in seen_classes), set(int(c) for c in unseen_classes)
        self.noise_dim, self.embed_dim, self.feature_dim = noise_dim, embed_dim, feature_dim
        self.batch_size_generator, self.real_seen_features = batch_size_generator, real_seen_features
        # RNG hooks (train_pascal_GMMN.py:216,229 draw on the CPU generator and move to the GPU)
        self.noise_fn = noise_fn or (lambda n: torch.rand((n, noise_dim)))
        self.index_fn = index_fn or (lambda n: torch.randint(low=0, high=n, size=(batch_size_generator,)))
        self.mask_fn = mask_fn  # optional: Dropout keep mask [n, hidden] for the generator (parity tests)

    def training_step(self, image, target, embedding, real_features=None):
        """image [B,3,H,W], target [B,H,W] float labels, embedding [B,E,H,W] per-pixel class embeddings (CUDA).
        Returns (classifier loss tensor, generator_loss_batch float, list of per-update generator losses)."""
        model = self.model.module if hasattr(self.model, "module") else self.model
        dev = image.device
        if real_features is None:
            with torch.no_grad():                                              # :154-157
                real_features = model.forward_before_class_prediction(image)
        fake_features = torch.zeros(real_features.shape, device=dev)           # :160-162
        generator_loss_batch, g_losses = 0.0, []
        fh, fw = real_features.shape[2], real_features.shape[3]
        for i, (rf, tg, emb) in enumerate(zip(real_features, target, embedding)):
            generator_loss_sample = 0.0
            rf = rf.permute(1, 2, 0).contiguous().view((-1, self.feature_dim))  # :170-174
            tg = nn.functional.interpolate(tg.view(1, 1, tg.shape[0], tg.shape[1]), size=(fh, fw),
                                           mode="nearest").view(-1)            # :175-179
            emb = nn.functional.interpolate(emb.view(1, *emb.shape), size=(fh, fw), mode="nearest")
            emb = emb.permute(0, 2, 3, 1).contiguous().view((-1, self.embed_dim))  # :180-195
            fake_i = torch.zeros(rf.shape, device=dev)
            unique_class = torch.unique(tg)                                    # :201
            has_unseen = any(int(u) in self.unseen for u in unique_class)      # :204-207
            for idx_in in unique_class:
                if idx_in != 255:
                    self.optimizer_generator.zero_grad()
                    idx_class = tg == idx_in
                    real_c, emb_c = rf[idx_class], emb[idx_class]
                    z = self.noise_fn(emb_c.shape[0]).to(dev)                  # :216-218
                    if self.mask_fn is not None:
                        fake_c = self.generator(emb_c, z.float(), keep_mask=self.mask_fn(emb_c.shape[0]).to(dev))
                    else:
                        fake_c = self.generator(emb_c, z.float())              # :220-222
                    if int(idx_in) in self.seen and not has_unseen:            # :224-227
                        ridx = self.index_fn(fake_c.shape[0]).to(dev)          # :229-233
                        g_loss = self.criterion_generator(fake_c[ridx], real_c[ridx])
                        g_losses.append(g_loss.item())
                        generator_loss_sample += g_losses[-1]
                        g_loss.backward()
                        self.optimizer_generator.step()                        # :239-240
                    fake_i[idx_class] = fake_c.detach().clone()                # :242
            generator_loss_batch += generator_loss_sample / len(unique_class)
            src = rf if (self.real_seen_features and not has_unseen) else fake_i   # :244-259
            fake_features[i] = src.view((fh, fw, self.feature_dim)).permute(2, 0, 1)
        self.optimizer.zero_grad()                                             # :261
        output = model.forward_class_prediction(fake_features.detach(), image.size()[2:])
        loss = self.criterion(output, target)
        loss.backward()
        self.optimizer.step()                                                  # :265-267
        return loss, generator_loss_batch, g_losses


class _PinnedRing:
    """Small host->device uploads (index lists of a step's plan) through a ring of pinned buffers: a pageable
    `torch.tensor(list).to(device)` blocks the host until the stream has drained (4 ms per step while the feature graph
    runs, tools/profile_config5_host.py); a pinned non_blocking copy is just enqueued.  A slot is reused only after the
    copy that last read it has executed (event)."""

    def __init__(self, slots=4, capacity=1 << 16):
        self.buf = [torch.empty(capacity, dtype=torch.int64).pin_memory() for _ in range(slots)]
        self.done = [None] * slots
        self.k = 0

    def upload(self, values, dev):
        """values: list of Python ints -> int64 device tensor"""
        n = len(values)
        if n > self.buf[0].numel():
            return torch.tensor(values, dtype=torch.int64).to(dev)
        k = self.k
        self.k = (k + 1) % len(self.buf)
        if self.done[k] is not None:
            self.done[k].synchronize()
        host = self.buf[k][:n]
        host.copy_(torch.tensor(values, dtype=torch.int64))
        out = host.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.done[k] = ev
        return out


class ZS3StepFused(ZS3Step):
    """Same iteration as `ZS3Step`, with the per-(image, class) generator updates executed by the fused
    work-list kernel (`zs3_gmmn_train_fused`, csrc/gmmn_fused.cu) instead of ~65 launches + 2 host syncs each.

    What changes relative to `ZS3Step` (the arithmetic of every update is the same; `tests/test_step2_gpu.py`
    holds both against the same oracle):
      * labels of the whole batch are down-sampled, histogrammed and stably sorted on the device once; ONE
        device->host copy (the [B, 256] class histogram) replaces `torch.unique` + a boolean mask per class;
      * the per-pixel embedding map is never down-sampled or gathered into [n_c, 300] matrices: an update reads
        the 128 embedding rows it needs straight from the full-resolution map through a row gather;
      * the MLP runs on the 128 sampled rows only (the rows the loss sees, train_pascal_GMMN.py:229-237); features
        for all pixels are generated only where the reference uses them (images holding an unseen class, or
        `real_seen_features=False`), with the weights the reference would have used at that point;
      * the sequential updates of consecutive images are queued and executed by one launch; the queue is flushed
        before anything that reads the generator weights;
      * `noise_fn=None` draws z ~ U[0,1) for the 128 sampled rows on the device (the reference draws n_c x 300
        values on the host and copies them, `:216-218`), `index_fn=None` draws the sampled row indices of all
        updates on the device in one call (the reference: one host `torch.randint` per class, `:229`); pass
        `noise_fn` / `index_fn` to reproduce the reference's host streams.  With both defaults and no injected
        mask, the work list is packed with a handful of numpy vector operations instead of per-update Python.
    """

    def __init__(self, *args, noise_fn=None, index_fn=None, graph_features=False, fuse_classifier_loss=True,
                 tensor_core_bulk=True, world_size=1, **kw):
        super().__init__(*args, noise_fn=noise_fn, index_fn=index_fn, **kw)
        # data parallel (one process per GPU, images sharded): ONE all-reduce per iteration over
        # [generator delta | pred_conv gradients], see parallel.exchange_step2
        self.world_size = int(world_size)
        self._device_index = index_fn is None     # default: sampled row indices drawn on the device for all updates
        from .gmmn_fused import FusedGeneratorUpdater
        self._device_noise = noise_fn is None
        self.fuse_classifier_loss = fuse_classifier_loss
        # features for whole unseen-class images: one image-level generator pass on the tcgen05 kernel (fp32x3)
        # instead of one fp32 SIMT GEMM pair per class
        self.tensor_core_bulk = tensor_core_bulk
        self.profile = None   # set to {} to collect per-segment (host ms, CUDA events) of the next training_step
        # graph_features: capture the (frozen-weight, no_grad) feature extraction in a CUDA graph on first use and
        # replay it afterwards -- ~350 launches per step issued by one graph launch instead of by the interpreter
        self.graph_features, self._feat_graph, self._feat_in, self._feat_out = graph_features, None, None, None
        # criterion_generator is GMMNLoss(...).build_loss(), a bound method of the loss object holding `sigma`
        sigma = getattr(getattr(self.criterion_generator, "__self__", None), "sigma", None) or (2, 5, 10, 20, 40, 80)
        self.updater = FusedGeneratorUpdater(self.generator, self.optimizer_generator, sigma=sigma)
        self._src_index = {}
        self._ring = None

    def _nearest_source_index(self, in_hw, out_hw, dev):
        """flat source-pixel index of every destination pixel under F.interpolate(mode='nearest') (`:175-195`),
        obtained from torch's own rule by interpolating an index image (exact: indices < 2**24)"""
        key = (tuple(in_hw), tuple(out_hw), str(dev))
        if key not in self._src_index:
            idx = torch.arange(in_hw[0] * in_hw[1], device=dev, dtype=torch.float32).view(1, 1, *in_hw)
            self._src_index[key] = nn.functional.interpolate(idx, size=out_hw, mode="nearest").view(-1).to(torch.int32)
        return self._src_index[key]

    def _extract_features(self, model, image):
        """model.forward_before_class_prediction(image) under no_grad (`:154-157`), optionally through a CUDA graph"""
        if not self.graph_features:
            with torch.no_grad():
                return model.forward_before_class_prediction(image)
        from . import functional as ZF
        dev = image.device
        if self._feat_graph is None or self._feat_in.shape != image.shape:
            if ZF._RngState.device_counter is None:
                ZF._RngState.device_counter = torch.zeros(1, dtype=torch.int64, device=dev)
            self._feat_in = image.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):          # lazy initialisations (function attributes, scratch) outside the capture
                    model.forward_before_class_prediction(self._feat_in)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self._feat_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._feat_graph), torch.no_grad():
                self._feat_out = model.forward_before_class_prediction(self._feat_in)
        ZF._RngState.device_counter.add_(1 << 32)  # fresh Dropout masks per replay
        self._feat_in.copy_(image, non_blocking=True)
        self._feat_graph.replay()
        return self._feat_out

    def _mark(self, name):
        if self.profile is not None:
            import time
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.profile.setdefault("marks", []).append((name, time.perf_counter(), ev))

    def profile_summary(self):
        """per-segment host and device milliseconds of the profiled step (call after torch.cuda.synchronize())"""
        marks = self.profile["marks"]
        return {b[0]: {"host_ms": (b[1] - a[1]) * 1e3, "gpu_ms": a[2].elapsed_time(b[2])} for a, b in zip(marks, marks[1:])}

    @torch.no_grad()
    def _generate_image(self, emb_map, src, tg_i, fh, fw, noise=None, keep_mask=None, table=None):
        """Generated features for EVERY pixel of one image (`:211-222,242`: the reference calls the generator once
        per class; rows are independent, so one call over all pixels with each pixel's own class embedding is the
        same map), as two 1x1 "convolutions" over the fh x fw grid on the tcgen05 implicit-GEMM kernel in fp32x3
        mode (zs3_b200/parity.py: three-way bf16 operand split, fp32 TMEM accumulation, ~2e-6 of an fp32 GEMM).
        Pixels labelled 255 keep zeros, as in the reference.  `noise` [hw, noise_dim] / `keep_mask` [hw, hidden] are
        injectable for tests (default: drawn on the device).  Returns fp32 [fh*fw, feature_dim]."""
        from . import gmmn_ops as G
        from . import kernels as K
        from . import parity as P
        upd = self.updater
        dev = src.device
        hw, kin = fh * fw, self.embed_dim + self.noise_dim
        x = torch.zeros((1, fh, fw, K.cpad(kin)), dtype=torch.float32, device=dev)
        xv = x.view(hw, -1)
        if table is None:
            xv[:, :self.embed_dim] = emb_map.reshape(self.embed_dim, -1)[:, src.long()].t()
        else:
            xv[:, :self.embed_dim] = self._table_rows(table, tg_i)
        xv[:, self.embed_dim:kin] = torch.rand((hw, self.noise_dim), device=dev) if noise is None else noise
        w1, b1, w2, b2 = (t.detach() for t in upd.params)
        h1 = P.conv_fp32([x], [kin], w1.view(upd.hidden, kin, 1, 1), 1, 1, 1, 0, 1, upd.hidden, bias=self._pad_bias(b1))
        hd = G.LeakyDropout.apply(h1.view(hw, -1), upd.act.negative_slope, upd.drop.p, self.generator.training, keep_mask)
        y = P.conv_fp32([hd.view(1, fh, fw, -1)], [upd.hidden], w2.view(upd.feat, upd.hidden, 1, 1), 1, 1, 1, 0, 1,
                        upd.feat, bias=self._pad_bias(b2))
        out = y.view(hw, -1)[:, :upd.feat]
        return torch.where((tg_i == 255)[:, None], torch.zeros_like(out), out)

    @staticmethod
    def _table_rows(table, labels):
        """E_table[label] as the reference's dataloader builds the per-pixel map (dataloaders/datasets/base.py:45-51):
        ignore pixels (255) carry the embedding of class 0"""
        lab = labels.long()
        return table[torch.where((lab < 0) | (lab >= table.shape[0]), torch.zeros_like(lab), lab)].contiguous()

    @staticmethod
    def _pad_bias(b):
        from . import kernels as K
        n = K.cpad(b.numel())
        return b if n == b.numel() else torch.cat([b, b.new_zeros(n - b.numel())])

    def _extra_classifier_backward(self, model, state):
        """hook between the classifier's backward and optimizer.step() (used by ZS3StepGCN)"""

    def _generator_params(self):
        """every generator parameter this runner trains (their per-iteration delta is averaged over the ranks)"""
        return list(self.updater.params)

    def _classifier_loss(self, model, features, image, target):
        """criterion(forward_class_prediction(features, input_size), target).  When the criterion is this package's
        cross entropy, the x4 bilinear upsample is evaluated inside the loss kernels from the low-resolution class
        scores (SegmentationLosses.UpsampledCrossEntropyLoss: same value, the 354 MB logits are never written)."""
        from .utils.loss import SegmentationLosses
        owner = getattr(self.criterion, "__self__", None)
        if (self.fuse_classifier_loss and isinstance(owner, SegmentationLosses)
                and getattr(self.criterion, "__func__", None) is SegmentationLosses.CrossEntropyLoss
                and tuple(target.shape[1:]) == tuple(image.shape[2:])
                and model.num_classes <= 64 and target.shape[-1] <= 640):   # limits of zs3_upsample_ce_*
            scores = model.decoder.forward_class_prediction(features)
            return owner.UpsampledCrossEntropyLoss(scores, model.num_classes, target)
        return self.criterion(model.forward_class_prediction(features, image.size()[2:]), target)

    @staticmethod
    def _feature_grid(h, w):
        """spatial size of forward_before_class_prediction's output: stem conv 7x7/2 pad 3, max-pool 3x3/2 pad 1
        (resnet.py:79-82); the decoder works at that low-level resolution (decoder.py:33-37)"""
        def down(n):
            return (n - 1) // 2 + 1
        return down(down(h)), down(down(w))

    def training_step(self, image, target, embedding=None, real_features=None, class_embeddings=None):
        """embedding: the reference's per-pixel map [B, E, H, W] (dataloaders/datasets/base.py:45-51 builds it as
        E_table[label]; 5.05 GB per bs=16 513x513 batch), OR class_embeddings: the [C, E] table itself -- every row the
        step needs is then gathered as table[label of the pixel] on the device and the map never exists (labels outside
        [0, C) -- 255 -- never reach the generator)."""
        from . import gmmn_fused as GF
        model = self.model.module if hasattr(self.model, "module") else self.model
        dev = image.device
        if (embedding is None) == (class_embeddings is None):
            raise ValueError("pass exactly one of embedding (per-pixel map) and class_embeddings ([C, E] table)")
        table = None if class_embeddings is None else class_embeddings.to(dev).float().contiguous()
        nb = image.shape[0]
        in_hw = tuple(target.shape[1:])
        mark = self._mark
        mark("start")
        fh, fw = self._feature_grid(*in_hw) if real_features is None else tuple(real_features.shape[2:])
        hw = fh * fw
        # ---- label work list first: its one host sync then overlaps nothing, and the feature extraction enqueued
        # right after it runs on the GPU while the host plans the updates below
        src = self._nearest_source_index(in_hw, (fh, fw), dev)                          # [hw] int32
        tg = target.reshape(nb, -1)[:, src.long()].long()                               # nearest down-sampling `:175-179`
        hist = torch.zeros((nb, 256), dtype=torch.int32, device=dev)
        hist.scatter_add_(1, tg.clamp(0, 255), torch.ones_like(tg, dtype=torch.int32))
        order = torch.argsort(tg, dim=1, stable=True).to(torch.int32)                   # raster order inside a class
        hist_h = hist.cpu().numpy()                                                     # the step's one label sync
        mark("labels")
        if real_features is None:
            real_features = self._extract_features(model, image)                       # `:154-157`
        mark("features")
        real_features = real_features.contiguous().float()
        if embedding is not None:
            embedding = embedding.contiguous().float()
        fd = real_features.shape[1]
        assert tuple(real_features.shape[2:]) == (fh, fw), (real_features.shape, fh, fw)

        # ---- host plan: the (image, class) visits in the reference's order, drawing the injected randomness in the
        # reference's order (noise, [mask], indices); entries are executed in this order below
        plan, n_unique, image_has_unseen = [], [], []
        for i in range(nb):
            classes = np.nonzero(hist_h[i])[0].tolist()                                  # == torch.unique (sorted)
            n_unique.append(len(classes))
            has_unseen = any(c in self.unseen for c in classes)
            image_has_unseen.append(has_unseen)
            need_fake = has_unseen or not self.real_seen_features
            # nothing injected and no update interleaved with the generation: one image-level generator call
            image_level = need_fake and has_unseen and self._device_noise and self.mask_fn is None and self.tensor_core_bulk
            if image_level:
                plan.append(("image", i, hw, 0, None, None, None))
            off = 0
            for c in classes:
                n_c, start = int(hist_h[i, c]), off
                off += n_c
                if c == 255:
                    continue
                z_full = None if self._device_noise else self.noise_fn(n_c)
                m_full = None if self.mask_fn is None else self.mask_fn(n_c)
                if need_fake and not image_level:
                    plan.append(("bulk", i, n_c, start, z_full, m_full, None))
                if c in self.seen and not has_unseen:
                    plan.append(("item", i, n_c, start, z_full, m_full, None if self._device_index else self.index_fn(n_c)))

        # ---- device side of all updates at once: sampled pixels in the feature grid and in the input grid, noise
        upd = [e for e in plan if e[0] == "item"]
        rows = self.batch_size_generator
        if upd:
            if self._ring is None and dev.type == "cuda":
                self._ring = _PinnedRing()
            up = (lambda v: self._ring.upload(v, dev)) if self._ring is not None else \
                (lambda v: torch.tensor(v, dtype=torch.int64).to(dev))
            # one upload for the three index lists of the plan: [base | n_c | image]
            packed = up([e[1] * hw + e[3] for e in upd] + [e[2] for e in upd] + [e[1] for e in upd])
            base, n_c_i, img_of = packed[:len(upd)], packed[len(upd):2 * len(upd)], packed[2 * len(upd):]
            if self._device_index:   # floor(u * n_c), u ~ U[0,1): uniform over the class's pixels, with replacement
                n_c_all = n_c_i.to(torch.float32)
                u = torch.rand((len(upd), rows), device=dev)
                ridx_all = torch.minimum((u * n_c_all[:, None]).floor(), n_c_all[:, None] - 1).to(torch.int32)
            else:
                ridx_all = torch.stack([e[6].to(torch.int32) for e in upd]).to(dev)      # [n, rows] one H2D
                rows = ridx_all.shape[1]
            pix_all = order.view(-1)[base[:, None] + ridx_all.long()].contiguous()      # [n, rows] feature-grid pixels
            if table is None:
                spix_all = src[pix_all.long()].contiguous()                              # same pixels, input grid
            else:   # rows of the class table: the label of each sampled pixel (constant per update)
                spix_all = tg[img_of[:, None], pix_all.long()].to(torch.int32).contiguous()
            z_all = torch.rand((len(upd), rows, self.noise_dim), device=dev) if self._device_noise else None

        mark("plan+index")
        gen_snapshot = None
        if self.world_size > 1:
            gen_snapshot = torch.cat([p.detach().reshape(-1) for p in self._generator_params()])
        fake_features = torch.zeros(real_features.shape, device=dev)
        fake_by_image = {}
        queue, keep, owners, loss_chunks = [], [], [e[1] for e in upd], []
        # nothing injected per update: the whole work list is packed with a handful of numpy vector operations
        vec = self._device_noise and self.mask_fn is None and bool(upd)
        if vec:
            arr = GF.pack_items_vectorized(
                images=np.array([e[1] for e in upd], dtype=np.int64), rows=rows,
                emb=((embedding.data_ptr(), embedding.stride(0) * 4, in_hw[0] * in_hw[1]) if table is None else
                     (table.data_ptr(), 0, 1, table.stride(0))), emb_rows=spix_all,
                noise=z_all, real=(real_features.data_ptr(), real_features.stride(0) * 4, hw), real_rows=pix_all,
                keep_rows=ridx_all)
        done = [0, 0]   # [updates launched, updates visited]

        def flush():
            if vec:
                if done[1] > done[0]:
                    loss_chunks.append(self.updater.run(arr[done[0]:done[1]], self.embed_dim, self.noise_dim,
                                                        keepalive=[spix_all, pix_all, ridx_all, z_all]))
                    done[0] = done[1]
            elif queue:
                loss_chunks.append(self.updater.run(list(queue), self.embed_dim, self.noise_dim, keepalive=list(keep)))
                queue.clear()
                keep.clear()

        for kind, i, n_c, start, z_full, m_full, _ in plan:
            z_dev = None if z_full is None else z_full.to(dev).float().contiguous()
            m_dev = None if m_full is None else m_full.to(dev).to(torch.uint8).contiguous()
            if kind == "image":
                flush()                                                                  # weights as of this point
                fake_by_image[i] = self._generate_image(None if table is not None else embedding[i], src, tg[i], fh, fw,
                                                        table=table)
                continue
            if kind == "bulk":
                flush()                                                                  # weights as of this point
                pix_c = order[i, start:start + n_c].long()
                if i not in fake_by_image:
                    fake_by_image[i] = torch.zeros((hw, fd), device=dev)
                with torch.no_grad():
                    if table is None:
                        emb_c = embedding[i].reshape(self.embed_dim, -1)[:, src[pix_c].long()].t().contiguous()
                    else:
                        emb_c = self._table_rows(table, tg[i][pix_c])
                    z_gen = z_dev if z_dev is not None else torch.rand((n_c, self.noise_dim), device=dev)
                    if m_dev is not None:
                        fake_c = self.generator(emb_c, z_gen, keep_mask=m_dev)
                    else:
                        fake_c = self.generator(emb_c, z_gen)                           # `:220-222`
                    fake_by_image[i][pix_c] = fake_c                                     # `:242`
                continue
            k = done[1]
            done[1] += 1
            if vec:
                continue
            ridx, pix, spix = ridx_all[k], pix_all[k], spix_all[k]
            emb_src = (GF.row_source(embedding[i], spix, row_stride=1, col_stride=in_hw[0] * in_hw[1]) if table is None
                       else GF.row_source(table, spix))
            noise_src = GF.row_source(z_all[k]) if z_dev is None else GF.row_source(z_dev, ridx)
            real_src = GF.row_source(real_features[i], pix, row_stride=1, col_stride=hw)
            queue.append(GF.pack_item(emb_src, noise_src, real_src, rows, keep_mask=m_dev, keep_rows=ridx))
            keep.extend([z_dev, m_dev])
        flush()
        mark("generator")
        for i in range(nb):                                                              # `:244-259`
            if self.real_seen_features and not image_has_unseen[i]:
                fake_features[i] = real_features[i]
            else:
                fake_features[i] = fake_by_image[i].view(fh, fw, fd).permute(2, 0, 1) if i in fake_by_image else 0
        self.optimizer.zero_grad()
        loss = self._classifier_loss(model, fake_features.detach(), image, target)      # `:261-264`
        loss.backward()
        self._extra_classifier_backward(model, dict(real_features=real_features, labels=tg, embedding=embedding,
                                                    table=table, src=src, grid=(fh, fw),
                                                    image_has_unseen=image_has_unseen))
        if self.world_size > 1:
            from .parallel import exchange_step2
            head = [p for g in self.optimizer.param_groups for p in g["params"] if p.grad is not None]
            exchange_step2(self._generator_params(), gen_snapshot, head, self.world_size)
        self.optimizer.step()
        mark("classifier")
        g_losses = torch.cat(loss_chunks).tolist() if loss_chunks else []
        per_image = [0.0] * nb
        for j, v in enumerate(g_losses):
            per_image[owners[j]] += v
        generator_loss_batch = sum(per_image[i] / n_unique[i] for i in range(nb))
        return loss, generator_loss_batch, g_losses


class ZS3StepGCN(ZS3StepFused):
    """Step-2 iteration of the GCN-context variant (zs3/train_context_GMMN_GCNcontext.py:270-460; BASELINE configs[4]):
    the ZS3Net iteration of `ZS3StepFused` plus, per image, the semantic-cluster graph of its label map, one update of
    the graph generator (`GMMNnetwork_GCN`, MMD between generated and real node features) and a cluster-level
    cross-entropy term on the classifier (`GCN_weight`).

    Host logic and arithmetic are checked on the CPU against the oracle (tests/test_kernel_emulation.py) and on the GPU
    by tests/test_step2_gpu.py::test_gcn_context_step_matches_oracle.
    Differences from the reference's host code: the cluster graphs of all images come from ONE launch of
    `zs3_label_components` (`:307-321` runs a Python DFS per image after three D2H copies), node embeddings / features
    are gathered on the device at the seed pixels, and the nodes of the batch are laid out on an [n/8, 8] grid (padded
    with ignore labels) for the classifier instead of [n, 1]."""

    def __init__(self, *args, generator_gcn=None, optimizer_generator_gcn=None, gcn_weight=0.1, gcn_noise_fn=None,
                 gcn_mask_fn=None, max_nodes=256, **kw):
        super().__init__(*args, **kw)
        if generator_gcn is None or optimizer_generator_gcn is None:
            raise ValueError("ZS3StepGCN needs generator_gcn and optimizer_generator_gcn")
        self.generator_gcn, self.optimizer_generator_gcn = generator_gcn, optimizer_generator_gcn
        self.gcn_weight, self.max_nodes = float(gcn_weight), max_nodes
        self.gcn_noise_fn = gcn_noise_fn      # optional: n -> [n, noise_dim] (tests); default torch.rand on the device
        self.gcn_mask_fn = gcn_mask_fn        # optional: n -> [n, hidden] Dropout keep mask of the graph generator (tests)
        self.last_gcn_losses = []
        # the graph-generator updates of a batch run as ONE work list of the fused kernel (items carry the adjacency
        # matrix) when the generator is this package's GMMNnetwork_GCN with a plain Adam; anything else (stand-in
        # modules in the CPU tests, other optimizers, > 128 nodes) takes the module path, one call at a time
        self.updater_gcn = None
        try:
            from .gmmn_fused import FusedGeneratorUpdater
            sigma = getattr(getattr(self.criterion_generator, "__self__", None), "sigma", None) or (2, 5, 10, 20, 40, 80)
            self.updater_gcn = FusedGeneratorUpdater(generator_gcn, optimizer_generator_gcn, sigma=sigma)
            if not self.updater_gcn.graph:
                self.updater_gcn = None
        except (NotImplementedError, ValueError, AttributeError, TypeError):
            self.updater_gcn = None

    def _generator_params(self):
        return list(self.updater.params) + list(self.generator_gcn.parameters())

    def _extra_classifier_backward(self, model, state):
        from . import graph as ZG
        real, labels, embedding, src = state["real_features"], state["labels"], state["embedding"], state["src"]
        fh, fw = state["grid"]
        nb, fd, hw = real.shape[0], real.shape[1], fh * fw
        dev = real.device
        n_nodes, node_label, node_seed, adj, _ = ZG.label_components(labels.float(), fh, fw, max_nodes=self.max_nodes)
        counts = n_nodes.tolist()                                                   # one sync for the batch
        feats, targets, self.last_gcn_losses = [], [], []
        from . import gmmn_fused as GF
        fused = self.updater_gcn is not None and real.is_cuda
        queue, keep, queued_updates = [], [], []

        def flush():
            if queue:
                losses = self.updater_gcn.run(list(queue), self.embed_dim, self.noise_dim, keepalive=list(keep))
                self.last_gcn_losses += [losses[k] for k in queued_updates]
                queue.clear()
                keep.clear()
                queued_updates.clear()

        for i, n in enumerate(counts):
            if n > self.max_nodes:
                raise RuntimeError(f"image {i}: {n} clusters exceed max_nodes={self.max_nodes}")
            if n <= 1:                                                              # adj_mat is None (`:93-97,323`)
                continue
            seeds = node_seed[i, :n].long()
            targets.append(node_label[i, :n].float())                               # `:323-324`
            if state.get("table") is None:
                emb_n = embedding[i].reshape(self.embed_dim, -1)[:, src[seeds].long()].t().contiguous()   # seed embeddings `:58`
            else:   # 255-regions are nodes too; the reference's map holds E[0] there (datasets/base.py:46-50)
                emb_n = self._table_rows(state["table"], labels[i][seeds])
            real_n = real[i].reshape(fd, -1)[:, seeds].t().contiguous()             # seed features `:72-74`
            z = (torch.rand((n, self.noise_dim), device=dev) if self.gcn_noise_fn is None
                 else self.gcn_noise_fn(n).to(dev).float())                         # `:404`
            has_unseen = state["image_has_unseen"][i]
            keep_real = self.real_seen_features and not has_unseen
            if fused and n <= GF.MAX_ROWS:
                # one item of the fused work list: forward (both graph convolutions), MMD loss, backward, Adam -- or the
                # forward alone for an image holding an unseen class (`:413-428`)
                adj_i = adj[i, :n, :n].contiguous()
                m8 = None if self.gcn_mask_fn is None else self.gcn_mask_fn(n).to(dev).to(torch.uint8).contiguous()
                out = None if keep_real else torch.empty((n, fd), dtype=torch.float32, device=dev)
                z = z.contiguous()
                queue.append(GF.pack_item(GF.row_source(emb_n), GF.row_source(z), GF.row_source(real_n), n, keep_mask=m8,
                                          adj=adj_i, out=out, forward_only=has_unseen))
                keep.extend([emb_n, z, real_n, adj_i, m8, out])
                if not has_unseen:
                    queued_updates.append(len(queue) - 1)
                feats.append(real_n if keep_real else out)
                continue
            flush()                                                                 # keep the updates in image order
            self.optimizer_generator_gcn.zero_grad()                                # `:402`
            if self.gcn_mask_fn is not None:
                fake_n = self.generator_gcn(emb_n, z, adj[i, :n, :n].contiguous(),
                                            keep_mask=self.gcn_mask_fn(n).to(dev).to(torch.uint8).contiguous())
            else:
                fake_n = self.generator_gcn(emb_n, z, adj[i, :n, :n].contiguous())  # `:407-409`
            if not has_unseen:                                                      # `:413-419`
                g_loss = self.criterion_generator(fake_n, real_n)
                g_loss.backward()
                self.optimizer_generator_gcn.step()
                self.last_gcn_losses.append(g_loss.detach())
            feats.append(real_n if keep_real else fake_n.detach())                  # `:421-428`
        flush()
        if not feats:
            return
        x = torch.cat(feats, 0)                                                     # [N, fd]
        t = torch.cat(targets, 0)
        pad = (-x.shape[0]) % 8
        if pad:
            x = torch.cat([x, x.new_zeros(pad, fd)], 0)
            t = torch.cat([t, t.new_full((pad,), 255.0)], 0)                        # ignored by the loss
        grid = (x.shape[0] // 8, 8)
        x = x.t().reshape(1, fd, *grid).contiguous()                                # `:439-443` ([1, fd, N, 1] there)
        out = model.forward_class_prediction(x, grid)                               # same-size "upsample" = identity
        loss_gcn = self.gcn_weight * self.criterion(out, t.view(1, *grid))          # `:444-453`
        loss_gcn.backward()
        self.last_gcn_cluster_loss = loss_gcn.detach()

