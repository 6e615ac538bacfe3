"""Import stub (pasco/maskpls/interpolate.py:3-4; only used by dead knn_up code)."""
def set_verbose(*a, **k):
    pass
