_default = ["numpy"]


def get_default_backend():
    return _default[0]


def set_default_backend(b):
    _default[0] = b if isinstance(b, str) else getattr(b, "name", "numpy")
