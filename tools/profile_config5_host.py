"""Host-side profile (cProfile) of one config-5 step (ZS3StepGCN): where the Python time of the classifier segment goes."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
orig = torch.cuda.synchronize
state = {}


def grab(*a, **k):
    return orig(*a, **k)


# reuse bench.config5_rate's construction by monkeypatching the step class to profile its last call
from zs3_b200 import step2 as S  # noqa: E402

real_step = S.ZS3StepGCN.training_step
calls = {"n": 0}


def wrapped(self, *a, **k):
    calls["n"] += 1
    if calls["n"] == 6:      # after warm-up
        pr = cProfile.Profile()
        pr.enable()
        out = real_step(self, *a, **k)
        torch.cuda.synchronize()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(45)
        return out
    return real_step(self, *a, **k)


S.ZS3StepGCN.training_step = wrapped
print(bench.config5_rate(dev, 8, 513, steps=5, warmup=3)["ms_per_step"])
