"""Drop-in for the one torch_scatter call on PaSCo's path (unet3d_sparse_v2.py:79)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from pasco_b200.me import scatter_max  # noqa: F401
