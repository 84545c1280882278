from . import xla_bridge  # noqa: F401
