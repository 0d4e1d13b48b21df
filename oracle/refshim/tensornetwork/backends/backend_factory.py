_BACKENDS = {}
_INSTANTIATED_BACKENDS = {}


def get_backend(backend):
    if not isinstance(backend, str):
        return backend
    if backend in _INSTANTIATED_BACKENDS:
        return _INSTANTIATED_BACKENDS[backend]
    _INSTANTIATED_BACKENDS[backend] = _BACKENDS[backend]()
    return _INSTANTIATED_BACKENDS[backend]
