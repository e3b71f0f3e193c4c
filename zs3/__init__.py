"""Import-path shim: `zs3.*` resolves to the B200-native implementation in `zs3_b200.*`, so the reference
trainers' imports (zs3/train_pascal_GMMN.py:9-18) work unchanged against this repository.

The reference's `zs3` is a namespace package (no __init__.py); with this repository placed BEFORE the reference
checkout on sys.path, the modules that are outside the hot path (zs3.dataloaders, zs3.parsing, zs3.utils.saver,
zs3.utils.lr_scheduler, ...) must still resolve to the reference's files: every other `zs3` directory found on
sys.path is appended to this package's search path (ours stays first, so the hot-path modules win)."""
import os
import sys


def _extend_search_path(path, parts):
    here = os.path.abspath(path[0])
    for entry in list(sys.path):
        cand = os.path.abspath(os.path.join(entry or ".", *parts))
        if os.path.isdir(cand) and cand != here and cand not in path:
            path.append(cand)


_extend_search_path(__path__, ("zs3",))
