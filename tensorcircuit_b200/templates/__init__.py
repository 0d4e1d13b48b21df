from . import measurements  # noqa: F401
from . import blocks  # noqa: F401
