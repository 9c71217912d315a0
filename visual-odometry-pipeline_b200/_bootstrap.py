"""Makes `import vo_b200` work for the drop-in modules, which the reference layout runs as top-level scripts
from this directory (`python3 vo_runner.py`, CWD-relative `config/vo_params.yaml`)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import vo_b200  # noqa: E402,F401
from vo_b200 import ops  # noqa: E402,F401
