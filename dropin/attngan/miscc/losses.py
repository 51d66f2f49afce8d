from mog_b200.attngan.miscc.losses import *  # noqa: F401,F403
