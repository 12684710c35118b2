"""Install the UNMODIFIED reference package into ``baseline/_ref`` -- TEST INFRASTRUCTURE ONLY.

``tests/test_dropin_reference.py`` runs the reference's own callers of the renderer
(``differentiable_renderer/scripts/experiments.py``, ``estimation/simple_setup.py::SDFPipeline``) on top
of this repository's library, and beside it on the reference's own CUDA extension.  The GPU box has no
``/root/reference``, so the package is installed once here, the way the task contract describes
(``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref``; the build needs a
writable source tree, so pip runs on a copy under /tmp).  ``baseline/_ref`` is git-ignored: no reference
source enters the history; the directory travels to the GPU box with the snapshot like the built
``.so`` files.  The trained VAE the reference ships as a test fixture
(``tests/initilization/vae_model/mug.pt`` + ``mug.yaml``) is copied next to it as
``baseline/_ref/fixtures/``.

Usage:  python oracle/install_reference.py      (no-op when already installed)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")
FIXTURES = os.path.join(TARGET, "fixtures")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE, "setup.py"))


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "sdfest", "estimation", "simple_setup.py")) \
        and os.path.isfile(os.path.join(FIXTURES, "mug.pt"))


def install(force: bool = False) -> str | None:
    if installed() and not force:
        return TARGET
    if not available():
        return TARGET if installed() else None
    scratch = tempfile.mkdtemp(prefix="sdfest_ref_install_")
    try:
        src = os.path.join(scratch, "src")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git"))
        os.makedirs(TARGET, exist_ok=True)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--upgrade", "--find-links", "/opt/wheelhouse", "--target", TARGET, src])
        os.makedirs(FIXTURES, exist_ok=True)
        for name in ("mug.pt", "mug.yaml"):
            shutil.copy2(os.path.join(REFERENCE, "tests", "initilization", "vae_model", name),
                         os.path.join(FIXTURES, name))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print("reference package:", install(force="--force" in sys.argv))
