"""Place the UNMODIFIED reference files the `--impl reference` bench arm drives under ``baseline/_ref/``.

    python baseline/install_reference.py            # build container only: /root/reference is mounted there

``baseline/_ref/`` is git-ignored (no reference source ever enters the history) but not gpurun-ignored, so the
copy travels to the GPU box with the snapshot.  The reference has no setup.py / pyproject (it is a script tree:
`pip install /root/reference` fails with "neither 'setup.py' nor 'pyproject.toml' found"), so the install is a
verbatim file copy of the one module the PN hot path lives in: ``src/models/modelPN.py`` (imports torch + numpy
only).  ``trainPNLow.py`` / ``trainPNHigh.py`` need IPython + matplotlib and ``modelML.py`` needs
torch_geometric / torch_scatter -- none of them is in this image -- so they are not installable.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
FILES = ["src/models/modelPN.py"]


def install(dst_root: str = os.path.join(ROOT, "baseline", "_ref")) -> bool:
    if not os.path.isdir(REFERENCE):
        return False
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            print(f"baseline/_ref/{rel}: sha256 {hashlib.sha256(f.read()).hexdigest()[:16]}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
