"""Stage the UNMODIFIED reference package under baseline/_ref (git-ignored, shipped to the GPU box by gpurun).

    python baseline/stage_ref.py            # pip install --no-index --target baseline/_ref <copy of /root/reference>

The reference is pure Python (package `mop`, pyproject name `mop-transformers`).  /root/reference is read-only and
setuptools wants to write build/ and *.egg-info next to the sources, so the install runs from a copy under /tmp.
Nothing from the reference is committed to this repository: baseline/_ref/ is listed in .gitignore.
Only the `mop` package is installed (pyproject `include = ["mop*"]`); the experiment scripts are not.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"


def staged() -> bool:
    return os.path.isfile(os.path.join(TARGET, "mop", "models", "attention_variants.py"))


def stage(force: bool = False) -> str:
    """Returns 'staged', 'already' or 'unavailable: <why>'."""
    if staged() and not force:
        return "already"
    if not os.path.isdir(SOURCE):
        return "unavailable: /root/reference not present on this machine"
    tmp = tempfile.mkdtemp(prefix="mop_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns("outputs", "results", "docs", "*.ipynb", "__pycache__"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return "unavailable: pip install failed: " + (r.stderr.strip().splitlines() or ["?"])[-1]
        return "staged"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def import_path() -> str | None:
    """Directory to put on sys.path to import the reference `mop` package, or None."""
    return TARGET if staged() else None


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
