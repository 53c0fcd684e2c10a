"""Copy the UNMODIFIED reference (its Python package `src/` and `cfgs/`) into baseline/_ref/ (git-ignored; it ships to
the GPU box with the gpurun snapshot, it is never committed).

    python oracle/install_ref.py            # build container only: needs /root/reference

baseline/_ref is what the reference arm of bench.py, the one-epoch AP/WLL comparison (tools/amp_delta.py) and the
drop-in integration test (tests/test_gpu_dropin_e2e.py: `python -m src.train` / `python -m src.infer` with this repo's
overlay in front of it on PYTHONPATH) import on the box, where /root/reference does not exist.  The reference has no
setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` (the base contract's recipe) does not
apply: the package is a plain namespace directory and a copy is the install.  TEST / BASELINE INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B200CTR_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF, "src")):
        if verbose:
            print(f"install_ref: {REF} not present (GPU box?) -- keeping whatever is in {DST}")
        return os.path.isdir(os.path.join(DST, "src"))
    os.makedirs(DST, exist_ok=True)
    for sub in ("src", "cfgs"):
        d = os.path.join(DST, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(REF, sub), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "* copy.py"))
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("Unmodified copy of /root/reference/{src,cfgs} made by oracle/install_ref.py (git-ignored).\n")
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"install_ref: {n} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
