"""Import-path shim: `zs3.*` resolves to the B200-native implementation in `zs3_b200.*`, so the reference
trainers' imports (zs3/train_pascal_GMMN.py:9-18) work unchanged against this repository."""
