"""Drop-in shadow of code/coco/attngan/model.py."""
from mog_b200.attngan.model import *  # noqa: F401,F403
from mog_b200.attngan.model import MAX_OBJECTS, stn  # noqa: F401
