"""Import shim: the package lives in `divortio-lz4_b200/` (a name Python cannot import directly)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "divortio-lz4_b200"))

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
