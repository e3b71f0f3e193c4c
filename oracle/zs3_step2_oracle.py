"""CPU oracle of one ZS3Net step-2 iteration -- TEST INFRASTRUCTURE (see oracle/zs3_oracle.py's header).

Restates zs3/train_pascal_GMMN.py:152-268 with the functional oracle: given the decoder features (real features),
labels, per-pixel embeddings and injected randomness, runs the sequential per-(image, class) generator updates
(Adam), assembles the classifier input and performs the pred_conv SGD update.  Returns the losses and the updated
parameter dictionaries.
"""
import torch
import torch.nn.functional as F

import zs3_oracle as O


def step2(st_deeplab, st_gen, real_features, target, embedding, input_size, seen, unseen, noise_fn, index_fn, mask_fn,
          class_weight, lr=0.07, lr_generator=2e-4, momentum=0.9, weight_decay=5e-4, real_seen_features=True,
          feature_dim=256, embed_dim=300, gcn=None):
    """gcn (optional, zs3/train_context_GMMN_GCNcontext.py:307-330,400-454): dict(state=GCN generator state dict,
    noise_fn, mask_fn, weight=GCN_weight) -- adds the per-image cluster-graph generator update and the cluster-level
    cross-entropy term of the classifier."""
    gen = {k: v.clone().requires_grad_(True) for k, v in st_gen.items()}
    adam = {k: [torch.zeros_like(v), torch.zeros_like(v)] for k, v in gen.items()}
    adam_step = 0
    fh, fw = real_features.shape[2], real_features.shape[3]
    fake_features = torch.zeros_like(real_features)
    g_losses, generator_loss_batch = [], 0.0
    if gcn is not None:
        import zs3_graph_oracle as GO
        gst = {k: v.clone().requires_grad_(True) for k, v in gcn["state"].items()}
        gadam = {k: [torch.zeros_like(v), torch.zeros_like(v)] for k, v in gst.items()}
        gcn_step, gcn_feats, gcn_targets, gcn_losses = 0, [], [], []
    for i in range(real_features.shape[0]):
        rf = real_features[i].permute(1, 2, 0).reshape(-1, feature_dim)                  # :170-174
        tg = F.interpolate(target[i][None, None], size=(fh, fw), mode="nearest").view(-1)  # :175-179
        emb = F.interpolate(embedding[i][None], size=(fh, fw), mode="nearest")[0].permute(1, 2, 0).reshape(-1, embed_dim)
        fake_i = torch.zeros_like(rf)
        uniq = torch.unique(tg)
        has_unseen = any(int(u) in unseen for u in uniq)
        sample_loss = 0.0
        for c in uniq:
            if c == 255:
                continue
            sel = tg == c
            real_c, emb_c = rf[sel], emb[sel]
            dev = emb_c.device                                 # (CPU in the tests; bench.py's library baseline runs it on cuda)
            z = noise_fn(emb_c.shape[0]).to(dev)               # :216-218 draws on the host, then .cuda()
            mask = mask_fn(emb_c.shape[0]).to(dev)
            fake_c = O.gmmn_forward(gen, emb_c, z.float(), training=True, keep_mask=mask)
            if int(c) in seen and not has_unseen:
                ridx = index_fn(fake_c.shape[0]).to(dev)
                loss = O.moment_loss(fake_c[ridx], real_c[ridx])
                g_losses.append(loss.item())
                sample_loss += loss.item()
                grads = torch.autograd.grad(loss, list(gen.values()))
                adam_step += 1
                with torch.no_grad():
                    for (k, p), g in zip(gen.items(), grads):
                        O.adam_step(p, g, adam[k][0], adam[k][1], adam_step, lr=lr_generator)
            fake_i[sel] = fake_c.detach()
        generator_loss_batch += sample_loss / len(uniq)
        src = rf if (real_seen_features and not has_unseen) else fake_i
        fake_features[i] = src.view(fh, fw, feature_dim).permute(2, 0, 1)
        if gcn is not None:                                                            # `:307-330,400-428`
            _, node_label, node_seed, adj = GO.cluster_graph(tg.view(fh, fw).numpy())
            n = len(node_label)
            if n > 1:
                seeds = torch.from_numpy(node_seed).long()
                gcn_targets.append(torch.from_numpy(node_label.astype("float32")))
                emb_n, real_n = emb[seeds], rf[seeds]
                z = gcn["noise_fn"](n)
                fake_n = O.gmmn_gcn_forward(gst, emb_n, z.float(), torch.from_numpy(adj), training=True,
                                            keep_mask=gcn["mask_fn"](n))
                if not has_unseen:
                    gl = O.moment_loss(fake_n, real_n)
                    gcn_losses.append(gl.item())
                    grads = torch.autograd.grad(gl, list(gst.values()))
                    gcn_step += 1
                    with torch.no_grad():
                        for (k, p), g in zip(gst.items(), grads):
                            O.adam_step(p, g, gadam[k][0], gadam[k][1], gcn_step, lr=lr_generator)
                gcn_feats.append(real_n if (real_seen_features and not has_unseen) else fake_n.detach())
    w = st_deeplab["decoder.pred_conv.weight"].clone().requires_grad_(True)
    b = st_deeplab["decoder.pred_conv.bias"].clone().requires_grad_(True)
    out = O.class_prediction({"decoder.pred_conv.weight": w, "decoder.pred_conv.bias": b}, fake_features, input_size)
    loss = O.cross_entropy(out, target, weight=class_weight)
    total = loss
    if gcn is not None and gcn_feats:                                                  # `:436-454`
        x = torch.cat(gcn_feats, 0)
        out_gcn = F.conv2d(x.t().reshape(1, feature_dim, -1, 1), w, b)
        tgt = torch.cat(gcn_targets, 0).view(1, -1, 1)
        total = loss + gcn["weight"] * O.cross_entropy(out_gcn, tgt, weight=class_weight)
    gw, gb = torch.autograd.grad(total, [w, b])
    with torch.no_grad():
        bufs = [None, None]
        O.sgd_step([w, b], [gw, gb], bufs, lr, momentum, weight_decay)
    res = {"loss": loss.item(), "g_losses": g_losses, "generator_loss_batch": generator_loss_batch,
           "pred_conv.weight": w.detach(), "pred_conv.bias": b.detach(),
           "generator": {k: v.detach() for k, v in gen.items()}, "fake_features": fake_features}
    if gcn is not None:
        res["gcn_generator"] = {k: v.detach() for k, v in gst.items()}
        res["gcn_losses"] = gcn_losses
    return res
