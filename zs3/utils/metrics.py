from zs3_b200.utils.metrics import Evaluator, Evaluator_seen_unseen  # noqa: F401
