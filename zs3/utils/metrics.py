from zs3_b200.utils.metrics import Evaluator  # noqa: F401
