"""Launches the round-1b kernels once each on their full-size workloads (for ncu and for event timing):
the fused generator-update kernel on a 16-update work list, the cluster-graph kernel on 16 label maps at 129x129,
and the argmax + confusion-matrix kernel on 16 x 21 x 513 x 513 logits.  Prints event timings as JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from zs3_b200 import gmmn_fused as GF  # noqa: E402
from zs3_b200.graph import label_components  # noqa: E402
from zs3_b200.modeling.gmmn import GMMNnetwork  # noqa: E402
from zs3_b200.utils.metrics import Evaluator  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1)


NCU = os.environ.get("ZS3_NCU") == "1"   # under ncu: one warm-up + one launch of each kernel


def timed(fn, reps=5):
    if NCU:
        reps = 1
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


res = {}
# fused generator updates: features as an NCHW map [256][129*129], embedding shared per update
gen = GMMNnetwork(300, 300, 256, 256).cuda().train()
upd = GF.FusedGeneratorUpdater(gen, torch.optim.Adam(gen.parameters(), lr=2e-4))
hw = 129 * 129
feats = torch.relu(torch.randn(16, 256, hw, device=dev))
table = torch.randn(21, 300, device=dev) * 0.06
items, keep = [], []
for k in range(16):
    pix = torch.randint(0, hw, (128,), device=dev, dtype=torch.int32)
    z = torch.rand(128, 300, device=dev)
    keep += [pix, z]
    items.append(GF.pack_item(GF.row_source(table[k:k + 1], row_stride=0), GF.row_source(z),
                              GF.row_source(feats[k], pix, row_stride=1, col_stride=hw), 128))
us = timed(lambda: upd.run(items, 300, 300, keepalive=keep))
res["gmmn_train_fused"] = {"updates": 16, "us_per_launch": us, "us_per_update": us / 16}
stamps = torch.zeros((16, 8), dtype=torch.int64, device=dev)
upd.run(items, 300, 300, keepalive=keep, phase_stamps=stamps)
torch.cuda.synchronize()
st = stamps.cpu().double()
names = ["P1 hidden layer (+gather of the next item)", "P2 output layer", "P3 pairwise kernel + loss partials",
         "P4 loss + dY", "P5 dH", "P6 weight gradients + Adam"]
res["gmmn_train_fused"]["phase_us_mean_over_updates_1_15"] = {
    n: float((st[1:, k + 1] - st[1:, k]).mean() / 1e3) for k, n in enumerate(names)}
res["gmmn_train_fused"]["update_us_from_stamps"] = float((st[1:, 6] - st[1:, 0]).mean() / 1e3)

# cluster graph: 16 Voronoi label maps at 129 x 129
rng = np.random.RandomState(3)
yy, xx = np.mgrid[0:129, 0:129]
seg = np.zeros((16, 129, 129), dtype=np.float32)
for b in range(16):
    k = rng.randint(5, 40)
    sites = rng.randint(0, 129, size=(k, 2))
    cls = rng.randint(0, 10, size=k)
    seg[b] = cls[((yy[..., None] - sites[:, 0]) ** 2 + (xx[..., None] - sites[:, 1]) ** 2).argmin(-1)]
segc = torch.from_numpy(seg).cuda().reshape(16, -1)
us = timed(lambda: label_components(segc, 129, 129, max_nodes=256))
res["label_components"] = {"images": 16, "us_per_launch": us}

# argmax + confusion matrix at the validation size
logits = torch.randn(16, 21, 513, 513, device=dev)
target = torch.randint(0, 21, (16, 513, 513), device=dev).float()
ev = Evaluator(21)
us = timed(lambda: ev.add_batch_logits(target, logits))
nbytes = logits.numel() * 4 + target.numel() * 4
res["argmax_confusion"] = {"us_per_launch": us, "algorithmic_bytes": nbytes, "achieved_GBps": nbytes / us * 1e-3}
print(json.dumps(res))
