class _Config:
    """``config.update("jax_enable_x64", True)`` (train.py:24) is a no-op: the B200 path is FP64 end to end."""
    values = {}

    def update(self, name, value):
        self.values[name] = value


config = _Config()
