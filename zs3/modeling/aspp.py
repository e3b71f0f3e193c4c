from zs3_b200.modeling.aspp import ASPP, _ASPPModule, build_aspp  # noqa: F401
