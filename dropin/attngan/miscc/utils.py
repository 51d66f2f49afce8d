from mog_b200.attngan.miscc.utils import *  # noqa: F401,F403
