"""Test-side alias of oracle/oracle_binding.py (the checker bindings)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from oracle_binding import *  # noqa: F401,F403,E402
from oracle_binding import _ptr, _pi16, _pi8, _pu8, _pf  # noqa: F401,E402
