"""The one place tests import the CPU oracle from (oracle/ is test infrastructure, never on the product path)."""
import os
import sys

_ORACLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
if _ORACLE not in sys.path:
    sys.path.insert(0, _ORACLE)

import dram_oracle as O  # noqa: E402,F401
