from mog_b200.attngan.GlobalAttention import *  # noqa: F401,F403
from mog_b200.attngan.GlobalAttention import GlobalAttentionGeneral, conv1x1  # noqa: F401
