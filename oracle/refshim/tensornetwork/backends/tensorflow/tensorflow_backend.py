from ..abstract_backend import AbstractBackend


class TensorFlowBackend(AbstractBackend):
    """placeholder: only the numpy backend is functional in the shim"""

    def __init__(self, *a, **k):
        raise ImportError("tensorflow backend is not available in the tensornetwork shim")
