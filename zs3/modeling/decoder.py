from zs3_b200.modeling.decoder import Decoder, build_decoder  # noqa: F401
