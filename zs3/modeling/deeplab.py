from zs3_b200.modeling.deeplab import *  # noqa: F401,F403
from zs3_b200.modeling.deeplab import DeepLab  # noqa: F401
