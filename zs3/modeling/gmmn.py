from zs3_b200.modeling.gmmn import GMMNnetwork, GMMNnetwork_GCN  # noqa: F401
