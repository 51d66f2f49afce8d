from mog_b200.attngan.trainer import condGANTrainer  # noqa: F401
