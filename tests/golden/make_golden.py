"""Generates tests/golden/*.npz by running the REAL reference (valeoai/ZS3 at /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference modules are imported from /root/reference with the repo root kept OFF sys.path (this repo
also ships a `zs3` package) and with a stub for the missing third-party `pygcn` (zs3/modeling/gmmn.py:2).
Weights come from oracle.init_*_state (seeded), loaded into the reference modules with load_state_dict, so
the committed vectors pin: reference(weights, input) == oracle(weights, input).
Dropout: the reference's nn.Dropout modules use torch's global RNG; the fixtures are generated with
p forced to 0 on those modules (train-mode vectors) or in eval mode, which is stated in each file's `note`.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REPO, "oracle"))

# pygcn stub: only needed so that `import zs3.modeling.gmmn` succeeds.
pygcn = types.ModuleType("pygcn")
layers = types.ModuleType("pygcn.layers")


class GraphConvolution(torch.nn.Module):
    def __init__(self, i, o, bias=True):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.zeros(i, o))
        self.bias = torch.nn.Parameter(torch.zeros(o))

    def forward(self, x, adj):
        return torch.spmm(adj, torch.mm(x, self.weight)) + self.bias


layers.GraphConvolution = GraphConvolution
pygcn.layers = layers
sys.modules["pygcn"] = pygcn
sys.modules["pygcn.layers"] = layers

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

import zs3_oracle as O  # noqa: E402
from zs3.modeling.deeplab import DeepLab  # noqa: E402  (the reference)
from zs3.modeling.gmmn import GMMNnetwork  # noqa: E402
from zs3.utils.loss import GMMNLoss, SegmentationLosses  # noqa: E402

assert os.path.abspath(sys.modules["zs3.modeling.deeplab"].__file__).startswith(REF)

torch.set_num_threads(8)


def sub(t, step=4):
    return t.detach()[..., ::step, ::step].contiguous().numpy().astype(np.float32)


def synth_labels(n, h, w, nclass, seed):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, nclass, (n, (h + 15) // 16, (w + 15) // 16), generator=g).float()
    lab = torch.nn.functional.interpolate(lab[:, None], size=(h, w), mode="nearest")[:, 0]
    ign = torch.rand(n, h, w, generator=g) < 0.02
    lab[ign] = 255
    return lab


def deeplab_golden():
    out = {}
    N, H = 2, 65
    x = torch.randn(N, 3, H, H, generator=torch.Generator().manual_seed(11))
    target = synth_labels(N, H, H, 21, 12)
    out["x_seed"] = np.array(11)
    out["target"] = target.numpy().astype(np.float32)
    for mode in ("eval", "train"):
        st = O.init_deeplab_state(seed=1, randomize_bn=(mode == "eval"))
        model = DeepLab(num_classes=21, output_stride=16, sync_bn=False, pretrained=False)
        model.load_state_dict(st)
        if mode == "eval":
            model.eval()
        else:
            model.train()
            model.aspp.dropout.p = 0.0
            model.decoder.last_conv[3].p = 0.0
            model.decoder.last_conv[7].p = 0.0
        taps = {}
        hooks = [
            model.backbone.layer1.register_forward_hook(lambda m, i, o: taps.__setitem__("low_level", o)),
            model.backbone.layer2.register_forward_hook(lambda m, i, o: taps.__setitem__("layer2", o)),
            model.backbone.layer3.register_forward_hook(lambda m, i, o: taps.__setitem__("layer3", o)),
            model.backbone.layer4.register_forward_hook(lambda m, i, o: taps.__setitem__("backbone", o)),
            model.aspp.register_forward_hook(lambda m, i, o: taps.__setitem__("aspp", o)),
        ]
        logits = model(x)
        feat = model.forward_before_class_prediction(x) if mode == "eval" else None
        for h in hooks:
            h.remove()
        out[f"{mode}_logits"] = sub(logits)
        for k in ("low_level", "layer2", "layer3", "backbone", "aspp"):
            out[f"{mode}_{k}"] = sub(taps[k], 2)[:, ::8]
        if feat is not None:
            out["eval_features"] = sub(feat, 2)[:, ::8]
        if mode == "train":
            crit = SegmentationLosses(weight=None, cuda=False).build_loss("ce")
            w = torch.ones(21)
            w[[15, 16, 17, 18, 19]] = 100.0
            crit_w = SegmentationLosses(weight=w, cuda=False).build_loss("ce")
            loss = crit(logits, target)
            out["train_loss"] = np.array(loss.item(), dtype=np.float64)
            out["train_loss_weighted"] = np.array(crit_w(logits.detach(), target).item(), dtype=np.float64)
            model.zero_grad()
            loss.backward()
            grads = dict(model.named_parameters())
            for k in ("decoder.pred_conv.weight", "decoder.pred_conv.bias", "decoder.last_conv.4.weight",
                      "decoder.conv1.weight", "aspp.conv1.weight", "aspp.aspp3.atrous_conv.weight",
                      "backbone.layer4.2.conv2.weight", "backbone.layer3.10.conv1.weight",
                      "backbone.layer2.0.downsample.0.weight", "backbone.layer1.0.conv2.weight",
                      "backbone.conv1.weight", "backbone.bn1.weight", "backbone.layer3.5.bn2.bias"):
                gk = grads[k].grad.detach().reshape(-1)
                out["grad/" + k] = gk[:: max(1, gk.numel() // 512)][:512].numpy().astype(np.float32)
                out["gradnorm/" + k] = np.array(gk.double().norm().item())
            sd = model.state_dict()
            out["train_running_mean/backbone.bn1"] = sd["backbone.bn1.running_mean"].numpy()
            out["train_running_var/backbone.layer3.22.bn3"] = sd["backbone.layer3.22.bn3.running_var"].numpy()[::8]
    out["note"] = np.array("reference DeepLab(num_classes=21, os=16, sync_bn=False); weights oracle.init_deeplab_state(1, "
                           "randomize_bn=eval); x=randn(2,3,65,65, seed 11); train-mode vectors with Dropout p=0; "
                           "logits stored [::4,::4], feature taps [:, ::8, ::2, ::2]")
    np.savez_compressed(os.path.join(HERE, "deeplab_small.npz"), **out)
    print("deeplab_small.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def gmmn_golden():
    out = {}
    g = torch.Generator().manual_seed(21)
    st = O.init_gmmn_state(seed=3)
    gen = GMMNnetwork(300, 300, 256, 256)
    gen.load_state_dict(st)
    gen.eval()  # Dropout off: deterministic
    emb = torch.randn(200, 300, generator=g) * 0.06
    z = torch.rand(200, 300, generator=g)
    fake = gen(emb, z)
    real = torch.relu(torch.randn(200, 256, generator=g))
    # inputs are regenerated by the tests from seed 21 in the same draw order (emb, z, real, idx)
    out["input_checksum"] = np.array([emb.double().sum().item(), z.double().sum().item(), real.double().sum().item()])
    out["fake_eval"] = fake.detach().numpy()[::2]
    crit = GMMNLoss(sigma=[2, 5, 10, 20, 40, 80], cuda=False).build_loss()
    idx = torch.randint(0, 200, (128,), generator=g)
    out["idx"] = idx.numpy()
    fk = fake[idx].detach().requires_grad_(True)
    loss = crit(fk, real[idx])
    loss.backward()
    out["mmd_loss"] = np.array(loss.item(), dtype=np.float64)
    out["mmd_grad_fake"] = fk.grad.numpy()
    # one generator training step in train mode is RNG dependent (Dropout) -> pinned through the oracle with
    # an injected mask in tests; here: eval-mode loss gradient wrt generator parameters
    gen.zero_grad()
    loss2 = crit(gen(emb, z)[idx], real[idx])
    loss2.backward()
    for k, p in gen.named_parameters():
        gk = p.grad.detach().reshape(-1)
        out["gen_grad/" + k] = gk[:: max(1, gk.numel() // 2048)][:2048].numpy().astype(np.float32)
        out["gen_gradnorm/" + k] = np.array(gk.double().norm().item())
    # single-layer variant (hidden_size=0, gmmn.py:33-34)
    st0 = O.init_gmmn_state(seed=4, hidden=0)
    gen0 = GMMNnetwork(300, 300, 0, 256)
    gen0.load_state_dict(st0)
    out["fake_linear"] = gen0(emb, z).detach().numpy()[::4]
    out["note"] = np.array("reference GMMNnetwork(300,300,256,256).eval() with oracle.init_gmmn_state(3); "
                           "GMMNLoss sigma=[2,5,10,20,40,80] on 128 sampled rows")
    np.savez_compressed(os.path.join(HERE, "gmmn.npz"), **out)
    print("gmmn.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    deeplab_golden()
    gmmn_golden()
