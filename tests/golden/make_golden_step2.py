"""Generates tests/golden/step2.npz by running the REAL `Trainer.training` of the reference
(/root/reference/zs3/train_pascal_GMMN.py:134-311) for one iteration on CPU.

Run in the build container only:  python tests/golden/make_golden_step2.py
The trainer module cannot be imported here (tensorboardX / matplotlib are missing), so the method is compiled from the
reference file's syntax tree -- executed from where it lies, not copied -- and bound to a mock `self` whose members are
the REAL reference objects wherever arithmetic happens:

    self.generator            zs3.modeling.gmmn.GMMNnetwork (real; pygcn stubbed for the import only), train mode
    self.criterion_generator  zs3.utils.loss.GMMNLoss(...).build_loss()                     (real)
    self.criterion            zs3.utils.loss.SegmentationLosses(weight).build_loss("ce")    (real)
    self.optimizer_generator  torch.optim.Adam(lr=2e-4)      self.optimizer  torch.optim.SGD(momentum .9, wd 5e-4)
    self.model.module         zs3.modeling.deeplab.DeepLab (real) -- forward_class_prediction is the real method;
                              forward_before_class_prediction returns a FIXED synthetic feature tensor (the loop, not
                              the backbone, is what this fixture pins; the backbone has its own fixtures)

Randomness: the loop draws `torch.rand` (noise) and `torch.randint` (sampled rows) from the global CPU generator and
the generator's nn.Dropout draws a mask.  The exec namespace gets a proxy for `torch` whose rand / randint come from
numpy's frozen `RandomState(seed_k)` streams (k = draw counter; reproducible anywhere), and a forward hook on the
Dropout module records the keep masks (stored bit-packed).  tests/test_oracle.py replays those into
oracle/zs3_step2_oracle.step2 and must reproduce the stored losses and updated weights.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_FILE = REF + "/zs3/train_pascal_GMMN.py"

sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REPO, "oracle"))

pygcn = types.ModuleType("pygcn")
layers = types.ModuleType("pygcn.layers")
layers.GraphConvolution = type("GraphConvolution", (torch.nn.Module,), {})
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

NOISE_SEED0, INDEX_SEED0 = 1000, 5000


def noise_draw(k, n, dim=300):
    """k-th noise draw of the iteration: numpy RandomState stream (frozen algorithm), float32 in [0, 1)"""
    return torch.from_numpy(np.random.RandomState(NOISE_SEED0 + k).random_sample((n, dim)).astype(np.float32))


def index_draw(k, n, rows=128):
    return torch.from_numpy(np.random.RandomState(INDEX_SEED0 + k).randint(0, n, size=(rows,)).astype(np.int64))


class TorchProxy:
    """`torch` as seen by the compiled method: rand / randint replaced by reproducible streams, the rest untouched"""

    def __init__(self):
        self.noise_calls, self.index_calls = 0, 0

    def __getattr__(self, name):
        return getattr(torch, name)

    def rand(self, size):
        out = noise_draw(self.noise_calls, size[0], size[1])
        self.noise_calls += 1
        return out

    def randint(self, low, high, size):
        assert low == 0
        out = index_draw(self.index_calls, high, size[0])
        self.index_calls += 1
        return out


def load_training_method(proxy):
    tree = ast.parse(open(REF_FILE).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Trainer")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "training")

    class Bar(list):
        def set_description(self, *_):
            pass

    ns = {"torch": proxy, "nn": torch.nn, "tqdm": lambda it: Bar(it)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF_FILE, "exec"), ns)
    return ns["training"]


def inputs():
    """the configuration tests/test_step2_gpu.py uses: 3 images of 65x65, image 1 holds unseen class 17"""
    B, HW, C = 3, 65, 21
    g = torch.Generator().manual_seed(4)
    lab = torch.zeros(B, HW, HW)
    for i, cls in enumerate([[0, 3, 7], [0, 17, 5], [2, 9]]):
        grid = torch.randint(0, len(cls), (4, 4), generator=g)
        lab[i] = torch.tensor(cls, dtype=torch.float32)[grid].repeat_interleave(17, 0).repeat_interleave(17, 1)[:HW, :HW]
    lab[:, :2, :] = 255
    table = torch.randn(C, 300, generator=torch.Generator().manual_seed(8)) * 0.06
    emb = table[lab.clamp(max=C - 1).long()].permute(0, 3, 1, 2).contiguous()      # dataloaders/datasets/base.py:45-51
    image = torch.randn(B, 3, HW, HW, generator=torch.Generator().manual_seed(1))
    feats = torch.relu(torch.randn(B, 256, 17, 17, generator=torch.Generator().manual_seed(9)))  # O(1) post-ReLU features
    return image, lab, emb, feats


def main():
    torch.manual_seed(1)   # the reference's default seed (parsing.py:18-20); only the generator's Dropout masks use it
    C = 21
    unseen = [15, 16, 17, 18, 19]
    seen = [c for c in range(C) if c not in unseen]
    image, target, embedding, feats = inputs()
    st = O.init_deeplab_state(seed=1, randomize_bn=True)
    gst = O.init_gmmn_state(seed=3)
    deeplab = DeepLab(num_classes=C, sync_bn=False, pretrained=False)
    deeplab.load_state_dict(st)
    deeplab.forward_before_class_prediction = lambda img: feats.clone()
    gen = GMMNnetwork(300, 300, 256, 256)
    gen.load_state_dict(gst)
    gen.train()
    masks = []
    drop = next(m for m in gen.modules() if isinstance(m, torch.nn.Dropout))
    drop.register_forward_hook(lambda mod, inp, out: masks.append((out != 0) | (inp[0] == 0)))
    cw = torch.ones(C)
    cw[unseen] = 100.0
    g_losses = []
    crit_g_real = GMMNLoss(sigma=[2, 5, 10, 20, 40, 80], cuda=False).build_loss()

    def crit_g(a, b):
        v = crit_g_real(a, b)
        g_losses.append(float(v.item()))
        return v

    class NS:
        pass

    class Wrapped(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.module = m

    self = NS()
    self.model = Wrapped(deeplab)
    self.generator = gen
    self.criterion = SegmentationLosses(weight=cw, cuda=False).build_loss("ce")
    self.criterion_generator = crit_g
    self.optimizer = torch.optim.SGD([{"params": deeplab.get_1x_lr_params(), "lr": 0.007},
                                      {"params": deeplab.get_10x_lr_params(), "lr": 0.07}], momentum=0.9, weight_decay=5e-4)
    self.optimizer_generator = torch.optim.Adam(gen.parameters(), lr=2e-4)
    self.scheduler = lambda *a: None
    self.best_pred = 0.0
    self.writer = NS()
    self.writer.add_scalar = lambda *a: None
    self.summary = NS()
    self.summary.visualize_image = lambda *a: None
    args = NS()
    args.cuda, args.feature_dim, args.embed_dim, args.noise_dim = False, 256, 300, 300
    args.unseen_classes_idx_metric, args.seen_classes_idx_metric = unseen, seen
    args.batch_size_generator, args.real_seen_features = 128, True
    args.dataset, args.no_val, args.batch_size = "pascal", False, 3
    self.args = args
    # 10 loader entries so that `i % (num_img_tr // 10)` is defined; nine of them are single-image batches, which the
    # loop skips (train_pascal_GMMN.py:140)
    real_sample = {"image": image, "label": target, "label_emb": embedding}
    skip = {"image": image[:1], "label": target[:1], "label_emb": embedding[:1]}
    self.train_loader = [real_sample] + [skip] * 9
    proxy = TorchProxy()
    training = load_training_method(proxy)
    w0 = deeplab.decoder.pred_conv.weight.detach().clone()
    training(self, 0, args)
    assert proxy.noise_calls == len(masks) and len(g_losses) == proxy.index_calls
    out = {
        "note": np.array("one iteration of the REAL Trainer.training (train_pascal_GMMN.py:134-311) on CPU; inputs from "
                         "inputs(); noise/index draws = numpy RandomState(1000+k)/(5000+k); masks bit-packed per draw"),
        "g_losses": np.array(g_losses, dtype=np.float64),
        "n_noise_draws": np.array(proxy.noise_calls), "n_index_draws": np.array(proxy.index_calls),
        "mask_rows": np.array([m.shape[0] for m in masks]),
        "masks_packed": np.packbits(torch.cat(masks, 0).numpy().astype(np.uint8), axis=1),
        "pred_conv.weight": deeplab.decoder.pred_conv.weight.detach().numpy(),
        "pred_conv.bias": deeplab.decoder.pred_conv.bias.detach().numpy(),
    }
    for k, v in gen.state_dict().items():     # weights sub-sampled 4x4 (the fixture stays small), biases whole, plus norms
        a = v.detach().numpy()
        out["generator/" + k] = a[::4, ::4] if a.ndim == 2 else a
        out["generator_norm/" + k] = np.array(np.linalg.norm(a.astype(np.float64)))
        out["generator_delta_norm/" + k] = np.array(np.linalg.norm((a - gst[k].numpy()).astype(np.float64)))
    assert (w0 - deeplab.decoder.pred_conv.weight).abs().max() > 0
    untouched = [k for k, p in deeplab.named_parameters() if "pred_conv" not in k and p.grad is not None]
    assert not untouched, untouched     # only pred_conv receives a gradient in step 2
    np.savez_compressed(os.path.join(HERE, "step2.npz"), **out)
    print("generator updates", len(g_losses), "noise draws", proxy.noise_calls, "g_losses", np.round(g_losses, 5))


if __name__ == "__main__":
    main()
