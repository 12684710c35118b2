"""Import the UNMODIFIED reference package (baseline/_ref, see oracle/install_reference.py) on top of a
chosen native renderer module -- TEST INFRASTRUCTURE ONLY.

The reference's ``sdf_renderer.py`` obtains its native module with
``torch.utils.cpp_extension.load(name="sdf_renderer_cpp", ...)`` at import (:21-28).  ``load_reference``
patches that one call to hand back either ``sdfest_b200.compat.sdf_renderer_cpp`` (this repository's
drop-in) or the reference's own extension compiled into ``oracle/_ref``; everything else the reference
imports and this image lacks (open3d, matplotlib, ffmpeg, skimage, yoco, cpas_toolbox, healpy, ...) --
none of it on the renderer's path -- resolves to inert stand-ins.
"""
from __future__ import annotations

import importlib
import os
import sys
from importlib.machinery import ModuleSpec
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SITE = os.path.join(ROOT, "baseline", "_ref")
FIXTURES = os.path.join(REF_SITE, "fixtures")
_STUBBED = ("open3d", "matplotlib", "ffmpeg", "skimage", "yoco", "cpas_toolbox", "healpy", "trimesh", "pyrender",
            "pynput", "torchinfo", "mesh_to_sdf", "PySide2")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SITE, "sdfest", "estimation", "simple_setup.py"))


class _StubFinder:
    roots: set = set()

    @classmethod
    def find_spec(cls, name, path=None, target=None):
        return ModuleSpec(name, cls) if name.split(".")[0] in cls.roots else None

    @staticmethod
    def create_module(spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__name__, m.__spec__, m.__loader__ = [], spec.name, spec, _StubFinder
        return m

    @staticmethod
    def exec_module(module):
        pass


def _install_stubs() -> None:
    for name in _STUBBED:
        if name in _StubFinder.roots:
            continue
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001  (missing, or present but broken in this image)
            _StubFinder.roots.add(name)
    if _StubFinder not in sys.meta_path:
        sys.meta_path.insert(0, _StubFinder)
    # what the reference's callers need from their stand-ins
    import numpy as np

    if not hasattr(np, "float"):
        np.float = float  # simple_renderer.py:271 (removed from numpy 1.24)
    if "matplotlib" in _StubFinder.roots:
        plt = importlib.import_module("matplotlib.pyplot")
        ax = mock.MagicMock(name="axes")
        plt.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), ((ax, ax), (ax, ax))) if a[:2] == (2, 2) \
            else (mock.MagicMock(), (ax, ax))
    if "yoco" in _StubFinder.roots:
        yoco = importlib.import_module("yoco")
        yoco.resolve_path = lambda path, search_paths=None: os.path.expanduser(path)
        yoco.load_config = lambda config, current_dict=None, **k: {**(current_dict or {}), **config}


def load_reference(native_module):
    """Fresh import of the reference's ``sdfest`` package bound to ``native_module`` (an object with the
    reference's ``forward`` / ``backward``).  Returns the top-level package; submodules are reached with
    ``importlib.import_module`` while the returned context is current -- call ``purge()`` before binding
    another native module."""
    if not available():
        raise RuntimeError("baseline/_ref is empty: run python oracle/install_reference.py")
    _install_stubs()
    purge()
    if REF_SITE not in sys.path:
        sys.path.insert(0, REF_SITE)
    import torch.utils.cpp_extension as ce

    with mock.patch.object(ce, "load", lambda *a, **k: native_module):
        pkg = importlib.import_module("sdfest")
        renderer = importlib.import_module("sdfest.differentiable_renderer.sdf_renderer")
    assert renderer.sdf_renderer_cpp is native_module
    return pkg


def purge() -> None:
    for k in [k for k in sys.modules if k == "sdfest" or k.startswith("sdfest.")]:
        del sys.modules[k]
