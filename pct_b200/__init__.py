"""Import alias: the package lives in `online-3d-bpp-pct_b200/` (not a valid Python identifier), this stub
extends its own search path to that directory so `import pct_b200` works from the repo root."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "online-3d-bpp-pct_b200"))
from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: E402
