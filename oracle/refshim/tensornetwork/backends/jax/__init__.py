from . import jax_backend  # noqa: F401
