"""Stand-in for opt_einsum (path finders only): a left-to-right pairwise path.  The contraction
*order* only changes floating-point summation order, not the result (see oracle/refshim/
tensornetwork/__init__.py for why this shim exists)."""

from . import paths  # noqa: F401
