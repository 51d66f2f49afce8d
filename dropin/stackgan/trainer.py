from mog_b200.stackgan.trainer import GANTrainer  # noqa: F401
