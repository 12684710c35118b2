"""Build the reference's own CUDA renderer extension into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  The product (``sdfest_b200``) never imports this.

The reference ships its hot path as two translation units
(``sdfest/differentiable_renderer/csrc/sdf_renderer.cpp`` and
``sdf_renderer_cuda.cu``) that it JIT-builds at import
(``sdfest/differentiable_renderer/sdf_renderer.py:21-28``).  They are compiled
here *from where they lie* under ``/root/reference`` with one mechanical
2-token API fix streamed through ``sed`` into a scratch directory under /tmp
(``position.type()`` -> ``position.scalar_type()`` on sdf_renderer_cuda.cu:494
and :535 -- torch 2.11 removed the implicit conversion the old spelling relied
on; no arithmetic changes).  No reference source is copied into the repo; the
only artefact kept is the binary ``oracle/_ref/sdf_renderer_cpp.so`` (git-ignored,
shipped to the GPU box by gpurun) which the ``-m gpu`` parity tests and
``bench.py`` use as the GPU-side checker / baseline.

Usage:  python oracle/build_ref.py            (≈5 min; no-op when up to date)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/sdfest/differentiable_renderer/csrc"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "sdf_renderer_cpp.so")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_CSRC, "sdf_renderer_cuda.cu"))


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Compile the reference extension; returns the .so path (None if no sources)."""
    if not available():
        return OUT_SO if os.path.isfile(OUT_SO) else None
    srcs = [os.path.join(REF_CSRC, f) for f in ("sdf_renderer.cpp", "sdf_renderer_cuda.cu")]
    if (
        not force
        and os.path.isfile(OUT_SO)
        and all(os.path.getmtime(OUT_SO) >= os.path.getmtime(s) for s in srcs)
    ):
        return OUT_SO
    os.makedirs(OUT_DIR, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="sdfest_ref_build_")
    try:
        patched_cu = os.path.join(scratch, "sdf_renderer_cuda.cu")
        with open(patched_cu, "w") as f:
            subprocess.check_call(
                ["sed", "-e", "s/position\\.type()/position.scalar_type()/", srcs[1]], stdout=f
            )
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
        os.environ.setdefault("MAX_JOBS", "4")
        from torch.utils.cpp_extension import load

        build_dir = os.path.join(scratch, "build")
        os.makedirs(build_dir)
        load(
            name="sdf_renderer_cpp",
            sources=[srcs[0], patched_cu],
            build_directory=build_dir,
            extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"],
            verbose=verbose,
            is_python_module=False,  # just build; importing needs no GPU but keep it lazy
        )
        shutil.copy2(os.path.join(build_dir, "sdf_renderer_cpp.so"), OUT_SO)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return OUT_SO


def load_module():
    """Import the prebuilt reference extension (``forward``/``backward``), or None."""
    if not os.path.isfile(OUT_SO):
        return None
    import importlib.util

    import torch  # noqa: F401  (the .so links against libtorch)

    spec = importlib.util.spec_from_file_location("sdf_renderer_cpp", OUT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    so = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("reference extension:", so)
