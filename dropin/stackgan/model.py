"""Shim: put this directory first on PYTHONPATH inside code/coco/stackgan to run the reference main.py on libmog."""
from mog_b200.stackgan.model import *  # noqa: F401,F403
