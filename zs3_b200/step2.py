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
