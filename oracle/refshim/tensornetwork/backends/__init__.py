from . import abstract_backend, backend_factory  # noqa: F401
from . import numpy, jax, tensorflow, pytorch  # noqa: F401

base_backend = abstract_backend
