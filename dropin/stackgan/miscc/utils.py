from mog_b200.stackgan.miscc.utils import *  # noqa: F401,F403
