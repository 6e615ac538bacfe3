"""Test-only shim for the single torch_scatter call on the path (unet3d_sparse_v2.py:79)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from me_oracle import scatter_max  # noqa: F401
