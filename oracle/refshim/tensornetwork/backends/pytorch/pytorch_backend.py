from ..abstract_backend import AbstractBackend


class PyTorchBackend(AbstractBackend):
    """placeholder: only the numpy backend is functional in the shim"""

    def __init__(self, *a, **k):
        raise ImportError("pytorch backend is not available in the tensornetwork shim")
