from . import pytorch_backend  # noqa: F401
