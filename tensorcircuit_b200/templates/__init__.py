from . import measurements  # noqa: F401
from . import blocks  # noqa: F401
from . import graphs  # noqa: F401
