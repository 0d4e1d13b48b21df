from . import numpy_backend  # noqa: F401
