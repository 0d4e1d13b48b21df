from . import tensorflow_backend  # noqa: F401
