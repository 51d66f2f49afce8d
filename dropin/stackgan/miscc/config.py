from mog_b200.stackgan.miscc.config import *  # noqa: F401,F403
from mog_b200.stackgan.miscc.config import cfg, cfg_from_file  # noqa: F401
