"""ZS3Net step-2 training iteration (generator + classifier) on the B200 modules.

Follows the body of Trainer.training in zs3/train_pascal_GMMN.py:152-268 (identical in train_context_GMMN.py):
frozen-backbone feature extraction, per-(image, class) generator updates on 128 sampled pixels with the MMD
loss, fake/real feature assembly, classifier (pred_conv) update.  The reference trainer can drive the same
modules unchanged; this class is the in-repo runner used by tests and benchmarks.
"""
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
        values on the host and copies them, `:216-218`); pass `noise_fn` to reproduce the reference's stream.
    """

    def __init__(self, *args, noise_fn=None, **kw):
        super().__init__(*args, noise_fn=noise_fn, **kw)
        from .gmmn_fused import FusedGeneratorUpdater
        self._device_noise = noise_fn is None
        # criterion_generator is GMMNLoss(...).build_loss(), a bound method of the loss object holding `sigma`
        sigma = getattr(getattr(self.criterion_generator, "__self__", None), "sigma", None) or (2, 5, 10, 20, 40, 80)
        self.updater = FusedGeneratorUpdater(self.generator, self.optimizer_generator, sigma=sigma)
        self._src_index = {}

    def _nearest_source_index(self, in_hw, out_hw, dev):
        """flat source-pixel index of every destination pixel under F.interpolate(mode='nearest') (`:175-195`),
        obtained from torch's own rule by interpolating an index image (exact: indices < 2**24)"""
        key = (tuple(in_hw), tuple(out_hw), str(dev))
        if key not in self._src_index:
            idx = torch.arange(in_hw[0] * in_hw[1], device=dev, dtype=torch.float32).view(1, 1, *in_hw)
            self._src_index[key] = nn.functional.interpolate(idx, size=out_hw, mode="nearest").view(-1).to(torch.int32)
        return self._src_index[key]

    def training_step(self, image, target, embedding, real_features=None):
        from . import gmmn_fused as GF
        model = self.model.module if hasattr(self.model, "module") else self.model
        dev = image.device
        if real_features is None:
            with torch.no_grad():
                real_features = model.forward_before_class_prediction(image)
        real_features = real_features.contiguous().float()
        embedding = embedding.contiguous().float()
        nb, fd, fh, fw = real_features.shape
        hw = fh * fw
        in_hw = target.shape[1:]
        src = self._nearest_source_index(in_hw, (fh, fw), dev)                          # [hw] int32
        tg = target.reshape(nb, -1)[:, src.long()].long()                               # nearest down-sampling `:175-179`
        hist = torch.zeros((nb, 256), dtype=torch.int32, device=dev)
        hist.scatter_add_(1, tg.clamp(0, 255), torch.ones_like(tg, dtype=torch.int32))
        order = torch.argsort(tg, dim=1, stable=True).to(torch.int32)                   # raster order inside a class
        hist_h = hist.cpu().numpy()                                                     # the step's one label sync
        fake_features = torch.zeros(real_features.shape, device=dev)
        queue, keep, owners = [], [], []          # owners[k] = image index of queued / executed update k
        loss_chunks = []

        def flush():
            if queue:
                loss_chunks.append(self.updater.run(list(queue), self.embed_dim, self.noise_dim, keepalive=list(keep)))
                queue.clear()
                keep.clear()

        n_unique = []
        for i in range(nb):
            classes = [c for c in range(256) if hist_h[i, c] > 0]                       # == torch.unique (sorted)
            n_unique.append(len(classes))
            starts = {}
            off = 0
            for c in classes:
                starts[c] = off
                off += int(hist_h[i, c])
            has_unseen = any(c in self.unseen for c in classes)
            need_fake = has_unseen or not self.real_seen_features
            fake_i = torch.zeros((hw, fd), device=dev) if need_fake else None
            for c in classes:
                if c == 255:
                    continue
                n_c = int(hist_h[i, c])
                pix_c = order[i, starts[c]:starts[c] + n_c]                              # pixels of the class, raster order
                z_full = None if self._device_noise else self.noise_fn(n_c).to(dev).float().contiguous()
                m_full = None if self.mask_fn is None else self.mask_fn(n_c).to(dev).to(torch.uint8).contiguous()
                if need_fake:
                    flush()                                                              # weights as of this point
                    with torch.no_grad():
                        emb_c = embedding[i].reshape(self.embed_dim, -1)[:, src[pix_c.long()].long()].t().contiguous()
                        z_gen = z_full if z_full is not None else torch.rand((n_c, self.noise_dim), device=dev)
                        if m_full is not None:
                            fake_c = self.generator(emb_c, z_gen, keep_mask=m_full)
                        else:
                            fake_c = self.generator(emb_c, z_gen)
                        fake_i[pix_c.long()] = fake_c
                if c in self.seen and not has_unseen:
                    ridx = self.index_fn(n_c).to(dev).to(torch.int32).contiguous()
                    rows = int(ridx.numel())
                    pix = pix_c[ridx.long()].contiguous()                                # sampled pixels (feature grid)
                    spix = src[pix.long()].contiguous()                                  # same pixels in the input grid
                    emb_src = GF.row_source(embedding[i], spix, row_stride=1, col_stride=in_hw[0] * in_hw[1])
                    if z_full is not None:
                        noise_src = GF.row_source(z_full, ridx)
                    else:
                        z_full = torch.rand((rows, self.noise_dim), device=dev)
                        noise_src = GF.row_source(z_full)
                    real_src = GF.row_source(real_features[i], pix, row_stride=1, col_stride=hw)
                    queue.append(GF.pack_item(emb_src, noise_src, real_src, rows, keep_mask=m_full, keep_rows=ridx))
                    keep.extend([ridx, pix, spix, z_full, m_full])
                    owners.append(i)
            if self.real_seen_features and not has_unseen:
                fake_features[i] = real_features[i]
            else:
                fake_features[i] = fake_i.view(fh, fw, fd).permute(2, 0, 1)
        flush()
        self.optimizer.zero_grad()
        output = model.forward_class_prediction(fake_features.detach(), image.size()[2:])
        loss = self.criterion(output, target)
        loss.backward()
        self.optimizer.step()
        g_losses = torch.cat(loss_chunks).tolist() if loss_chunks else []
        per_image = [0.0] * nb
        for k, v in enumerate(g_losses):
            per_image[owners[k]] += v
        generator_loss_batch = sum(per_image[i] / n_unique[i] for i in range(nb))
        return loss, generator_loss_batch, g_losses
