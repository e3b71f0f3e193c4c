"""CPU oracle for the ZS3Net hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch functional restatement (torch CPU, fp32 or fp64) of what the reference computes on the
path BASELINE.json names.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under zs3_b200/ does.

Every function cites the reference lines it follows (paths relative to the reference checkout).
Parity pinning: tests/golden/make_golden.py runs the REAL reference modules (imported from
/root/reference in the build container) on seeded inputs with weights produced by `init_deeplab_state`
/ `init_gmmn_state` and stores sub-sampled outputs in tests/golden/*.npz; tests/test_oracle.py checks
this oracle against those files.  `graph_conv` (pygcn.layers.GraphConvolution, third-party, absent from
the reference tree and unpinned there) is "parity unpinned": it restates upstream pygcn's documented
`adj @ (x @ W) + b`.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default used everywhere in the reference
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------- structure
def resnet101_blocks(output_stride=16):
    """(name, inplanes, planes, stride, dilation, has_downsample) for the 33 bottlenecks.
    zs3/modeling/backbone/resnet.py:66-76 (strides/dilations), :121-184 (_make_layer/_make_MG_unit), :236."""
    if output_stride == 16:
        strides, dilations = [1, 2, 2, 1], [1, 1, 1, 2]
    elif output_stride == 8:
        strides, dilations = [1, 2, 1, 1], [1, 1, 2, 4]
    else:
        raise NotImplementedError
    layers = [3, 4, 23]
    mg = [1, 2, 4]
    out = []
    inplanes = 64
    for li, (planes, nblk) in enumerate(zip([64, 128, 256], layers)):
        for b in range(nblk):
            stride = strides[li] if b == 0 else 1
            ds = b == 0 and (stride != 1 or inplanes != planes * 4)
            out.append((f"backbone.layer{li + 1}.{b}", inplanes, planes, stride, dilations[li], ds))
            inplanes = planes * 4
    planes = 512
    for b in range(3):
        stride = strides[3] if b == 0 else 1
        ds = b == 0 and (stride != 1 or inplanes != planes * 4)
        out.append((f"backbone.layer4.{b}", inplanes, planes, stride, mg[b] * dilations[3], ds))
        inplanes = planes * 4
    return out


def aspp_dilations(output_stride=16):
    """zs3/modeling/aspp.py:47-52"""
    if output_stride == 16:
        return [1, 6, 12, 18]
    if output_stride == 8:
        return [1, 12, 24, 36]
    raise NotImplementedError


def _bn_names(prefix):
    return [f"{prefix}.weight", f"{prefix}.bias", f"{prefix}.running_mean", f"{prefix}.running_var",
            f"{prefix}.num_batches_tracked"]


def deeplab_param_shapes(num_classes=21, output_stride=16, global_avg_pool_bn=True):
    """Ordered {state_dict key: shape} of zs3.modeling.deeplab.DeepLab (680 entries at the defaults)."""
    shapes = {}

    def conv(name, cout, cin, k):
        shapes[name + ".weight"] = (cout, cin, k, k)

    def bn(name, c):
        shapes[name + ".weight"] = (c,)
        shapes[name + ".bias"] = (c,)
        shapes[name + ".running_mean"] = (c,)
        shapes[name + ".running_var"] = (c,)
        shapes[name + ".num_batches_tracked"] = ()

    conv("backbone.conv1", 64, 3, 7)
    bn("backbone.bn1", 64)
    for name, inpl, planes, stride, dil, ds in resnet101_blocks(output_stride):
        conv(name + ".conv1", planes, inpl, 1)
        bn(name + ".bn1", planes)
        conv(name + ".conv2", planes, planes, 3)
        bn(name + ".bn2", planes)
        conv(name + ".conv3", planes * 4, planes, 1)
        bn(name + ".bn3", planes * 4)
        if ds:
            conv(name + ".downsample.0", planes * 4, inpl, 1)
            bn(name + ".downsample.1", planes * 4)
    for i, k in enumerate([1, 3, 3, 3]):
        conv(f"aspp.aspp{i + 1}.atrous_conv", 256, 2048, k)
        bn(f"aspp.aspp{i + 1}.bn", 256)
    conv("aspp.global_avg_pool.1", 256, 2048, 1)
    if global_avg_pool_bn:
        bn("aspp.global_avg_pool.2", 256)
    conv("aspp.conv1", 256, 1280, 1)
    bn("aspp.bn1", 256)
    conv("decoder.conv1", 48, 256, 1)
    bn("decoder.bn1", 48)
    conv("decoder.last_conv.0", 256, 304, 3)
    bn("decoder.last_conv.1", 256)
    conv("decoder.last_conv.4", 256, 256, 3)
    bn("decoder.last_conv.5", 256)
    conv("decoder.pred_conv", num_classes, 256, 1)
    shapes["decoder.pred_conv.bias"] = (num_classes,)
    return shapes


def init_deeplab_state(seed=1, num_classes=21, output_stride=16, dtype=torch.float32, randomize_bn=False):
    """Deterministic random init following the reference's rules:
    backbone convs normal(0, sqrt(2/(k*k*Cout))) (resnet.py:199-204), ASPP/decoder convs kaiming_normal_
    fan_in (aspp.py:118-123, decoder.py:74-77), BN weight 1 / bias 0, pred_conv bias uniform(+-1/sqrt(fan_in))
    (nn.Conv2d default).  `randomize_bn` perturbs BN affine + running stats so eval-mode tests are not trivial."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for name, shape in deeplab_param_shapes(num_classes, output_stride).items():
        if name.endswith("num_batches_tracked"):
            st[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean"):
            st[name] = torch.randn(shape, generator=g) * 0.1 if randomize_bn else torch.zeros(shape)
        elif name.endswith("running_var"):
            st[name] = torch.rand(shape, generator=g) + 0.5 if randomize_bn else torch.ones(shape)
        elif len(shape) == 4:
            cout, cin, k, _ = shape
            if name.startswith("backbone."):
                std = math.sqrt(2.0 / (k * k * cout))
            else:
                std = math.sqrt(2.0 / (cin * k * k))
            st[name] = torch.randn(shape, generator=g) * std
        elif name == "decoder.pred_conv.bias":
            bound = 1.0 / math.sqrt(256)
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith(".weight"):
            st[name] = torch.rand(shape, generator=g) + 0.5 if randomize_bn else torch.ones(shape)
        else:
            st[name] = torch.randn(shape, generator=g) * 0.1 if randomize_bn else torch.zeros(shape)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in st.items()}


# --------------------------------------------------------------------------------------------- layers
def batch_norm(st, prefix, x, training, update_running=True):
    """F.batch_norm as called at zs3/modeling/sync_batchnorm/batchnorm.py:48-58 (single-device path)."""
    rm, rv = st[prefix + ".running_mean"], st[prefix + ".running_var"]
    if training and not update_running:
        rm, rv = rm.clone(), rv.clone()
    return F.batch_norm(x, rm, rv, st[prefix + ".weight"], st[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)


def dropout(x, p, training, masks, name):
    """nn.Dropout with an injectable keep-mask (masks[name], same shape as x, 1 = keep)."""
    if not training or p == 0.0:
        return x
    if masks == "torch":  # stock nn.Dropout (bench.py's library baseline: the reference's own GPU path, RNG not injected)
        return F.dropout(x, p, True)
    if masks is None or name not in masks:
        raise ValueError(f"oracle dropout '{name}' needs an explicit keep mask in training mode")
    return x * masks[name].to(x.dtype) / (1.0 - p)


def bottleneck(st, name, x, stride, dilation, has_ds, training):
    """zs3/modeling/backbone/resnet.py:33-53"""
    out = F.conv2d(x, st[name + ".conv1.weight"])
    out = F.relu(batch_norm(st, name + ".bn1", out, training))
    out = F.conv2d(out, st[name + ".conv2.weight"], stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(batch_norm(st, name + ".bn2", out, training))
    out = F.conv2d(out, st[name + ".conv3.weight"])
    out = batch_norm(st, name + ".bn3", out, training)
    if has_ds:
        res = F.conv2d(x, st[name + ".downsample.0.weight"], stride=stride)
        res = batch_norm(st, name + ".downsample.1", res, training)
    else:
        res = x
    return F.relu(out + res)


def backbone(st, x, training, output_stride=16, taps=None):
    """zs3/modeling/backbone/resnet.py:186-197: returns (x, low_level_feat)."""
    x = F.conv2d(x, st["backbone.conv1.weight"], stride=2, padding=3)
    x = F.relu(batch_norm(st, "backbone.bn1", x, training))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["stem"] = x
    low = None
    for name, inpl, planes, stride, dil, ds in resnet101_blocks(output_stride):
        x = bottleneck(st, name, x, stride, dil, ds, training)
        if taps is not None:
            taps[name] = x
        if name == "backbone.layer1.2":
            low = x
    return x, low


def aspp(st, x, training, output_stride=16, masks=None, drop_p=0.5, global_avg_pool_bn=True):
    """zs3/modeling/aspp.py:103-116"""
    dil = aspp_dilations(output_stride)
    outs = []
    for i in range(4):
        k = st[f"aspp.aspp{i + 1}.atrous_conv.weight"]
        pad = 0 if i == 0 else dil[i]
        y = F.conv2d(x, k, padding=pad, dilation=dil[i])
        outs.append(F.relu(batch_norm(st, f"aspp.aspp{i + 1}.bn", y, training)))
    g = F.adaptive_avg_pool2d(x, (1, 1))
    g = F.conv2d(g, st["aspp.global_avg_pool.1.weight"])
    if global_avg_pool_bn:
        g = batch_norm(st, "aspp.global_avg_pool.2", g, training)
    g = F.relu(g)
    g = F.interpolate(g, size=x.shape[2:], mode="bilinear", align_corners=True)
    y = torch.cat(outs + [g], dim=1)
    y = F.conv2d(y, st["aspp.conv1.weight"])
    y = F.relu(batch_norm(st, "aspp.bn1", y, training))
    return dropout(y, drop_p, training, masks, "aspp.dropout")


def decoder_features(st, x, low, training, masks=None, drop_p=(0.5, 0.1)):
    """zs3/modeling/decoder.py:42-52 (forward_before_class_prediction)"""
    low = F.conv2d(low, st["decoder.conv1.weight"])
    low = F.relu(batch_norm(st, "decoder.bn1", low, training))
    x = F.interpolate(x, size=low.shape[2:], mode="bilinear", align_corners=True)
    x = torch.cat((x, low), dim=1)
    x = F.conv2d(x, st["decoder.last_conv.0.weight"], padding=1)
    x = F.relu(batch_norm(st, "decoder.last_conv.1", x, training))
    x = dropout(x, drop_p[0], training, masks, "decoder.dropout0")
    x = F.conv2d(x, st["decoder.last_conv.4.weight"], padding=1)
    x = F.relu(batch_norm(st, "decoder.last_conv.5", x, training))
    return dropout(x, drop_p[1], training, masks, "decoder.dropout1")


def class_prediction(st, feat, input_size):
    """zs3/modeling/decoder.py:66-68 + zs3/modeling/deeplab.py:53-56"""
    y = F.conv2d(feat, st["decoder.pred_conv.weight"], st["decoder.pred_conv.bias"])
    return F.interpolate(y, size=input_size, mode="bilinear", align_corners=True)


def deeplab_forward(st, x, training=False, output_stride=16, masks=None, drop_p=(0.5, 0.5, 0.1), taps=None):
    """zs3/modeling/deeplab.py:40-45.  drop_p = (aspp, decoder0, decoder1); pass zeros to disable."""
    f, low = backbone(st, x, training, output_stride, taps)
    a = aspp(st, f, training, output_stride, masks, drop_p[0])
    feat = decoder_features(st, a, low, training, masks, drop_p[1:])
    logits = class_prediction(st, feat, x.shape[2:])
    if taps is not None:
        taps.update({"backbone": f, "low_level": low, "aspp": a, "features": feat})
    return logits


# --------------------------------------------------------------------------------------------- losses
def cross_entropy(logit, target, weight=None, ignore_index=255, batch_average=True):
    """zs3/utils/loss.py:31-46: weighted mean CE with ignore_index, then divided by the batch size."""
    n = logit.shape[0]
    loss = F.cross_entropy(logit, target.long(), weight=weight, ignore_index=ignore_index, reduction="mean")
    return loss / n if batch_average else loss


def moment_loss(gen_samples, x, sigma=(2, 5, 10, 20, 40, 80)):
    """zs3/utils/loss.py:92-115 (GMMNLoss.moment_loss incl. get_scale_matrix's [+1/N]*N ++ [-1/M]*M quirk)."""
    X = torch.cat((gen_samples, x), 0)
    XX = X @ X.t()
    X2 = torch.sum(X * X, 1, keepdim=True)
    exp = XX - 0.5 * X2 - 0.5 * X2.t()
    M, N = gen_samples.shape[0], x.shape[0]
    s = torch.cat((torch.ones(N, 1, dtype=X.dtype) / N, -torch.ones(M, 1, dtype=X.dtype) / M), 0).to(X.device)  # loss.py:93-96 builds it on the host, then .cuda()
    S = s @ s.t()
    loss = 0
    for v in sigma:
        loss = loss + torch.sum(S * torch.exp(exp / v))
    return torch.sqrt(loss)


# ----------------------------------------------------------------------------------------------- GMMN
def init_gmmn_state(seed=1, noise_dim=300, embed_dim=300, hidden=256, feat=256, dtype=torch.float32):
    """zs3/modeling/gmmn.py:23-35: xavier_uniform weights, bias 0.01; keys follow nn.Sequential indices."""
    g = torch.Generator().manual_seed(seed)

    def xavier(o, i):
        b = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * b

    if hidden:
        st = {"model.0.weight": xavier(hidden, noise_dim + embed_dim), "model.0.bias": torch.full((hidden,), 0.01),
              "model.3.weight": xavier(feat, hidden), "model.3.bias": torch.full((feat,), 0.01)}
    else:
        st = {"model.weight": xavier(feat, noise_dim + embed_dim), "model.bias": torch.full((feat,), 0.01)}
    return {k: v.to(dtype) for k, v in st.items()}


def gmmn_forward(st, embd, noise, training=False, keep_mask=None, drop_p=0.5):
    """zs3/modeling/gmmn.py:43-49 with the block of :17-21 (Linear -> LeakyReLU(0.2) -> Dropout(0.5) -> Linear)."""
    x = torch.cat((embd, noise), 1)
    if "model.weight" in st:
        return F.linear(x, st["model.weight"], st["model.bias"])
    h = F.leaky_relu(F.linear(x, st["model.0.weight"], st["model.0.bias"]), 0.2)
    if training and drop_p > 0:
        if keep_mask is None:
            raise ValueError("oracle GMMN dropout needs an explicit keep mask in training mode")
        h = h * keep_mask.to(h.dtype) / (1.0 - drop_p)
    return F.linear(h, st["model.3.weight"], st["model.3.bias"])


def graph_conv(x, adj, weight, bias):
    """pygcn.layers.GraphConvolution.forward (third-party tkipf/pygcn, NOT in the reference tree; parity unpinned):
    output = adj @ (x @ W) + b with W [in, out]."""
    return adj @ (x @ weight) + bias


def gmmn_gcn_forward(st, embd, noise, adj, training=False, keep_mask=None, drop_p=0.5):
    """zs3/modeling/gmmn.py:64-67"""
    x = graph_conv(torch.cat((embd, noise), 1), adj, st["gcn1.weight"], st["gcn1.bias"])
    x = F.leaky_relu(x, 0.2)
    if training and drop_p > 0:
        x = x * keep_mask.to(x.dtype) / (1.0 - drop_p)
    return graph_conv(x, adj, st["gcn2.weight"], st["gcn2.bias"])


# ------------------------------------------------------------------------------------------ optimizers
def sgd_step(params, grads, bufs, lr, momentum=0.9, weight_decay=5e-4, nesterov=False):
    """torch.optim.SGD semantics used at zs3/train_pascal.py:55-60 (first step: buf = grad)."""
    for i, (p, g) in enumerate(zip(params, grads)):
        d = g + weight_decay * p
        if bufs[i] is None:
            bufs[i] = d.clone()
        else:
            bufs[i].mul_(momentum).add_(d)
        d = d + momentum * bufs[i] if nesterov else bufs[i]
        p.sub_(lr * d)


def adam_step(p, g, m, v, step, lr=2e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (zs3/train_pascal_GMMN.py:65-67), in place; returns nothing."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, (v.sqrt() / math.sqrt(bc2)).add_(eps), value=-lr / bc1)


# ------------------------------------------------------------------------------------------- metrics
def confusion_matrix(gt, pred, num_class):
    """zs3/utils/metrics.py:73-77 (Evaluator._generate_matrix): rows = ground truth, columns = prediction; pixels
    whose label lies outside [0, num_class) (255 = ignore) are skipped.  numpy int64 [C, C]."""
    import numpy as np
    gt, pred = np.asarray(gt), np.asarray(pred)
    keep = (gt >= 0) & (gt < num_class)
    idx = num_class * gt[keep].astype(np.int64) + pred[keep].astype(np.int64)
    return np.bincount(idx, minlength=num_class * num_class).reshape(num_class, num_class)


def evaluator_scores(cm, seen=None, unseen=None):
    """zs3/utils/metrics.py:11-71: (pixel acc, class acc, mIoU, fwIoU) overall, and over seen / unseen rows."""
    import numpy as np
    cm = np.asarray(cm, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.diag(cm)
        acc_c = d / cm.sum(1)
        iou = d / (cm.sum(1) + cm.sum(0) - d)
        freq = cm.sum(1) / cm.sum()

        def scores(idx):
            rows = slice(None) if idx is None else idx
            f, i = freq[rows], iou[rows]
            return (d[rows].sum() / cm[rows, :].sum(), np.nanmean(np.nan_to_num(acc_c[rows])),
                    np.nanmean(np.nan_to_num(i)), (f[f > 0] * i[f > 0]).sum())
        out = {"all": scores(None)}
        if seen and unseen:
            out["seen"], out["unseen"] = scores(seen), scores(unseen)
    return out
