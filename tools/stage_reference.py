#!/usr/bin/env python
"""Stages the UNMODIFIED reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored, so it travels to the GPU box).

    python tools/stage_reference.py            # build container only: needs /root/reference

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>` — the
   reference's setup.py uses find_packages() but `pasco/` has no __init__.py, so the wheel it builds is EMPTY
   (dist-info only).
2. Therefore the package directory `pasco/` and `scripts/` are placed next to the dist-info verbatim — what the
   reference's documented `pip install -ve .` (editable: the source tree itself on sys.path) gives a user.
Nothing under baseline/_ref is ever committed; the product never imports it (only tests/test_dropin_reference.py and
bench.py --model reference-on-shim do, and they skip cleanly when it is absent).
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(SRC):
        raise SystemExit(f"{SRC} not present (GPU box?) — nothing to stage")
    tmp = "/tmp/_pasco_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp, ignore=shutil.ignore_patterns("*.pth", "*.gif", "teaser", "logs"))
    os.makedirs(DST, exist_ok=True)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", DST, "--upgrade", tmp], capture_output=True, text=True)
    print(r.stdout[-400:], r.stderr[-400:])
    if not os.path.isdir(os.path.join(DST, "pasco")):
        print("wheel was empty (namespace package): placing pasco/ and scripts/ verbatim")
        for d in ("pasco", "scripts"):
            shutil.rmtree(os.path.join(DST, d), ignore_errors=True)
            shutil.copytree(os.path.join(tmp, d), os.path.join(DST, d), ignore=shutil.ignore_patterns("__pycache__"))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f"staged {n} files under {DST}")


if __name__ == "__main__":
    main()
