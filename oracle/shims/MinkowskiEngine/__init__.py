"""Test-only import shim: lets the UNMODIFIED reference (`import MinkowskiEngine as ME`)
run on the CPU oracle inside the build container to generate tests/golden fixtures.
Never on the product path (product shim: compat/MinkowskiEngine → pasco_b200.me)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from me_oracle import *            # noqa: F401,F403
from me_oracle import utils, __version__, CoordinateManager, CoordinateMapKey  # noqa: F401
