"""numpy backend base class: the methods of tensornetwork's NumPyBackend that the reference's
NumpyBackend(numpy_backend.NumPyBackend, ExtendedBackend) inherits and the hot path calls."""

import numpy as np
import scipy.linalg

from ..abstract_backend import AbstractBackend


class NumPyBackend(AbstractBackend):
    def __init__(self):
        super().__init__()
        self.name = "numpy"

    def tensordot(self, a, b, axes):
        return np.tensordot(a, b, axes)

    def reshape(self, tensor, shape):
        return np.reshape(tensor, np.asarray(shape).astype(np.int32))

    def transpose(self, tensor, perm=None):
        return np.transpose(tensor, perm)

    def shape_concat(self, values, axis):
        return np.concatenate(values, axis)

    def slice(self, tensor, start_indices, slice_sizes):
        obj = tuple(slice(s, s + n) for s, n in zip(start_indices, slice_sizes))
        return tensor[obj]

    def shape_tensor(self, tensor):
        return tensor.shape

    def shape_tuple(self, tensor):
        return tensor.shape

    def sparse_shape(self, tensor):
        return self.shape_tensor(tensor)

    def shape_prod(self, values):
        return np.prod(values)

    def sqrt(self, tensor):
        return np.sqrt(tensor)

    def convert_to_tensor(self, tensor):
        return np.asarray(tensor)

    def outer_product(self, tensor1, tensor2):
        return np.tensordot(tensor1, tensor2, 0)

    def einsum(self, expression, *tensors, optimize=True):
        return np.einsum(expression, *tensors, optimize=optimize)

    def norm(self, tensor):
        return np.linalg.norm(tensor)

    def eye(self, N, dtype=None, M=None):
        return np.eye(N, M=M, dtype=dtype or np.float64)

    def ones(self, shape, dtype=None):
        return np.ones(shape, dtype=dtype or np.float64)

    def zeros(self, shape, dtype=None):
        return np.zeros(shape, dtype=dtype or np.float64)

    def conj(self, tensor):
        return np.conj(tensor)

    def eigh(self, matrix):
        return np.linalg.eigh(matrix)

    def trace(self, tensor, offset=0, axis1=-2, axis2=-1):
        return np.trace(tensor, offset=offset, axis1=axis1, axis2=axis2)

    def diagonal(self, tensor, offset=0, axis1=-2, axis2=-1):
        return np.diagonal(tensor, offset=offset, axis1=axis1, axis2=axis2)

    def diagflat(self, tensor, k=0):
        return np.diagflat(tensor, k=k)

    def abs(self, tensor):
        return np.abs(tensor)

    def sign(self, tensor):
        return np.sign(tensor)

    def addition(self, a, b):
        return a + b

    def subtraction(self, a, b):
        return a - b

    def multiply(self, a, b):
        return a * b

    def divide(self, a, b):
        return a / b

    def inv(self, matrix):
        return np.linalg.inv(matrix)

    def sin(self, tensor):
        return np.sin(tensor)

    def cos(self, tensor):
        return np.cos(tensor)

    def exp(self, tensor):
        return np.exp(tensor)

    def log(self, tensor):
        return np.log(tensor)

    def expm(self, matrix):
        return scipy.linalg.expm(matrix)

    def jit(self, fun, *args, **kwargs):
        return fun

    def sum(self, tensor, axis=None, keepdims=False):
        return np.sum(tensor, axis=axis, keepdims=keepdims)

    def matmul(self, a, b):
        return np.matmul(a, b)

    def item(self, tensor):
        return tensor.item()

    def power(self, a, b):
        return np.power(a, b)

    def eps(self, dtype):
        return np.finfo(dtype).eps
