"""Import stub: pasco/models/utils.py:7 imports h5py at module level but the hot path never uses it."""
def __getattr__(name):
    raise ImportError("h5py is not installed (stub from pasco_b200/compat)")
