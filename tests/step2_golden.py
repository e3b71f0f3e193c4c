"""Shared by the CPU oracle test and the GPU step-2 tests: rebuilds the inputs of tests/golden/make_golden_step2.py and
replays the recorded randomness of the REAL reference `Trainer.training` iteration stored in tests/golden/step2.npz
(noise / index draws are numpy RandomState streams, the generator's Dropout keep masks are stored bit-packed)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step2.npz")
NOISE_SEED0, INDEX_SEED0 = 1000, 5000
C, UNSEEN = 21, [15, 16, 17, 18, 19]
SEEN = [c for c in range(C) if c not in UNSEEN]


def inputs():
    B, HW = 3, 65
    g = torch.Generator().manual_seed(4)
    lab = torch.zeros(B, HW, HW)
    for i, cls in enumerate([[0, 3, 7], [0, 17, 5], [2, 9]]):
        grid = torch.randint(0, len(cls), (4, 4), generator=g)
        lab[i] = torch.tensor(cls, dtype=torch.float32)[grid].repeat_interleave(17, 0).repeat_interleave(17, 1)[:HW, :HW]
    lab[:, :2, :] = 255
    table = torch.randn(C, 300, generator=torch.Generator().manual_seed(8)) * 0.06
    emb = table[lab.clamp(max=C - 1).long()].permute(0, 3, 1, 2).contiguous()
    image = torch.randn(B, 3, HW, HW, generator=torch.Generator().manual_seed(1))
    feats = torch.relu(torch.randn(B, 256, 17, 17, generator=torch.Generator().manual_seed(9)))
    return image, lab, emb, feats, table


class Replay:
    """noise_fn / index_fn / mask_fn in the order the reference's loop drew them"""

    def __init__(self):
        self.gold = np.load(GOLD)
        rows = self.gold["mask_rows"]
        bits = np.unpackbits(self.gold["masks_packed"], axis=1)[:, :256].astype(bool)
        self.masks, off = [], 0
        for r in rows:
            self.masks.append(torch.from_numpy(bits[off:off + r].copy()))
            off += r
        self.reset()

    def reset(self):
        self.nk = self.ik = self.mk = 0

    def noise(self, n):
        out = torch.from_numpy(np.random.RandomState(NOISE_SEED0 + self.nk).random_sample((n, 300)).astype(np.float32))
        self.nk += 1
        return out

    def index(self, n):
        out = torch.from_numpy(np.random.RandomState(INDEX_SEED0 + self.ik).randint(0, n, size=(128,)).astype(np.int64))
        self.ik += 1
        return out

    def mask(self, n):
        m = self.masks[self.mk]
        assert m.shape[0] == n, (m.shape, n)
        self.mk += 1
        return m
