"""North-star tolerance: logits within 1e-3 (relative L2) of the reference on identical weights and inputs.
The bf16 throughput path cannot meet it (bf16 activation storage, DESIGN.md "Numerics"); the fp32x3 parity mode --
the same tcgen05 conv kernel fed with 3-way bf16-split operands as six K-segments, fp32 activations -- does."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

pytestmark = pytest.mark.gpu
TOL = 1e-3  # BASELINE.json north_star


def test_fp32x3_conv_kernel_is_fp32_grade():
    from zs3_b200 import kernels as K
    from zs3_b200 import parity as P
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 256, 33, 33, generator=g).cuda()
    w = (torch.randn(256, 256, 3, 3, generator=g) * 0.02).cuda()
    ref = F.conv2d(x.double(), w.double(), padding=2, dilation=2)
    xh = torch.empty(2, 33, 33, 256, device="cuda")
    xh.copy_(x.permute(0, 2, 3, 1))
    y = P.conv_fp32([xh], [256], w, 3, 3, 1, 2, 2, 256)
    e32 = rel_l2(F.conv2d(x, w, padding=2, dilation=2).double(), ref)
    e = rel_l2(y.permute(0, 3, 1, 2).double(), ref)
    print(f"fp32x3 conv rel_l2 vs fp64 = {e:.2e} (torch fp32 conv: {e32:.2e})")
    assert e < 5e-6  # fp32-grade: two orders of magnitude below the TF32 cuDNN path the reference uses on GPUs
    hi, mid, lo = P.split3(x)
    assert (hi.double() + mid.double() + lo.double() - x.double()).abs().max() <= 2 ** -22 * x.abs().max()


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_deeplab_logits_within_north_star_tolerance(mode):
    import zs3_oracle as O
    from zs3_b200 import parity as P
    from zs3_b200.modeling.deeplab import DeepLab
    x = torch.randn(2, 3, 65, 65, generator=torch.Generator().manual_seed(11))
    st = O.init_deeplab_state(seed=1, randomize_bn=(mode == "eval"))
    st64 = {k: (v.double() if v.is_floating_point() else v) for k, v in st.items()}
    taps = {}
    with torch.no_grad():
        ref = O.deeplab_forward(st, x, training=(mode == "train"), drop_p=(0.0, 0.0, 0.0), taps=taps)
        ref64 = O.deeplab_forward(st64, x.double(), training=(mode == "train"), drop_p=(0.0, 0.0, 0.0))
    model = DeepLab(num_classes=21, sync_bn=True, pretrained=False)
    model.load_state_dict(st)
    model = model.cuda()
    model.train() if mode == "train" else model.eval()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    out = P.deeplab_forward_fp32x3(model, x.cuda())
    feat = P.deeplab_forward_fp32x3(model, x.cuda(), return_features=True)
    e = rel_l2(out.cpu(), ref)
    e64 = rel_l2(out.cpu().double(), ref64)
    self_noise = rel_l2(ref.double(), ref64)
    print(f"{mode}: fp32x3 logits rel_l2 vs oracle fp32 = {e:.2e}, vs oracle fp64 = {e64:.2e}; "
          f"oracle fp32-vs-fp64 self noise = {self_noise:.2e}; features {rel_l2(feat.cpu(), taps['features']):.2e}")
    assert out.shape == ref.shape
    # the fp64 oracle is the ground truth both fp32 computations approximate
    assert e64 < TOL
    assert e < TOL or e64 <= 2 * self_noise


def test_config0_full_resolution_eval_forward():
    """BASELINE.json configs[0]: DeepLabv3+ ResNet-101 forward on 1x3x513x513 (the reference's CPU-runnable case):
    bf16 throughput path within its storage bound, fp32x3 parity mode within the north-star 1e-3."""
    import zs3_oracle as O
    from zs3_b200 import parity as P
    from zs3_b200.modeling.deeplab import DeepLab
    x = torch.randn(1, 3, 513, 513, generator=torch.Generator().manual_seed(1))
    st = O.init_deeplab_state(seed=1, randomize_bn=True)
    torch.set_num_threads(max(1, min(os.cpu_count() or 1, 32)))
    with torch.no_grad():
        ref = O.deeplab_forward(st, x, training=False)
    model = DeepLab(num_classes=21, sync_bn=False, pretrained=False)
    model.load_state_dict(st)
    model = model.cuda().eval()
    with torch.no_grad():
        fast = model(x.cuda())
    exact = P.deeplab_forward_fp32x3(model, x.cuda())
    e_fast, e_exact = rel_l2(fast.cpu(), ref), rel_l2(exact.cpu(), ref)
    print(f"config0 513x513: bf16 path rel_l2={e_fast:.2e}, fp32x3 rel_l2={e_exact:.2e}")
    assert tuple(fast.shape) == (1, 21, 513, 513)
    assert e_fast < 3e-2
    assert e_exact < TOL
