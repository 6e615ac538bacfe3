"""Import stub (pasco/data/semantic_kitti/io_data.py:7; real-data I/O only)."""
def __getattr__(name):
    raise ImportError("imageio is not installed (stub from pasco_b200/compat)")
