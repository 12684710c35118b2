"""Drop-in for the reference's pybind11 module ``sdf_renderer_cpp``.

The reference JIT-builds that module at import (``sdfest/differentiable_renderer/sdf_renderer.py:21-28``)
from ``csrc/sdf_renderer.cpp`` (module definition :88-91, ``forward`` :42-61, ``backward`` :63-86) and
``csrc/sdf_renderer_cuda.cu``.  This module has the same two functions with the same positional
arguments and return values, backed by ``libsdfrender.so`` through the C ABI (``sdfr_forward`` /
``sdfr_backward``, ``include/sdfrender.h``).  A sdfest maintainer replaces the ``load(...)`` call by

    from sdfest_b200.compat import sdf_renderer_cpp

and nothing else in the reference changes: ``SDFRendererFunctionGPU.forward/backward``
(sdf_renderer.py:311, :347) call ``sdf_renderer_cpp.forward / .backward`` exactly as before
(``tests/test_dropin_reference.py`` runs the reference's own callers that way).

Differences from the reference module, all inside its contract: outputs are allocated with
``torch.empty`` and fully written by the kernels (the reference needs ``torch::zeros``, cu:484,
525-528); launches go to the current CUDA stream of the tensors' device (the reference uses the legacy
default stream, cu:495, 536); any grid resolution works (the reference kernels hard-code 64, cu:225-230).
"""
from __future__ import annotations

from typing import List

import torch

from .. import _lib
from ..differentiable_renderer.sdf_renderer import _on_device_of, _stream


def _check_input(t: torch.Tensor, name: str) -> None:
    """CHECK_INPUT of sdf_renderer.cpp:9-13."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32 (the float64 instantiation of the reference never worked, cu:484-500)")


def forward(sdf: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor, inv_scale: torch.Tensor,
            width: int, height: int, cx: float, cy: float, fx: float, fy: float,
            threshold: float) -> List[torch.Tensor]:
    """sdf_renderer_forward (sdf_renderer.cpp:42-61): returns ``[depth_image (H,W)]``."""
    for t, n in ((sdf, "sdf"), (position, "position"), (orientation, "orientation"), (inv_scale, "inv_scale")):
        _check_input(t, n)
    with _on_device_of(sdf):  # OptionalCUDAGuard, cpp:58 (free when sdf lives on the current device)
        depth = torch.empty((int(height), int(width)), dtype=torch.float32, device=sdf.device)
        _lib.check(_lib.lib().sdfr_forward(
            sdf.data_ptr(), int(sdf.shape[-1]), 0, _lib.LAYOUT_DENSE, position.data_ptr(), orientation.data_ptr(),
            inv_scale.data_ptr(), 1, int(width), int(height), float(cx), float(cy), float(fx), float(fy),
            float(threshold), depth.data_ptr(), None, _stream()), "sdfr_forward")
    return [depth]


def backward(grad_depth_image: torch.Tensor, depth_image: torch.Tensor, sdf: torch.Tensor,
             position: torch.Tensor, orientation: torch.Tensor, inv_scale: torch.Tensor, width: int, height: int,
             cx: float, cy: float, fx: float, fy: float) -> List[torch.Tensor]:
    """sdf_renderer_backward (sdf_renderer.cpp:63-86): returns ``[grad_sdf, grad_position,
    grad_orientation, grad_inv_scale]`` shaped like the inputs (``zeros_like``, cu:525-528)."""
    grad_depth_image = grad_depth_image.contiguous()
    for t, n in ((grad_depth_image, "grad_depth_image"), (depth_image, "depth_image"), (sdf, "sdf"),
                 (position, "position"), (orientation, "orientation"), (inv_scale, "inv_scale")):
        _check_input(t, n)
    with _on_device_of(sdf):  # cpp:82
        g_sdf = torch.empty_like(sdf)  # cleared by the library (SDFR_ZERO_GRADS)
        exact = position.numel() == 3 and orientation.numel() == 4 and inv_scale.numel() == 1
        alloc = torch.empty_like if exact else torch.zeros_like  # over-long pose tensors: the tail reads 0
        g_p, g_q, g_is = alloc(position), alloc(orientation), alloc(inv_scale)
        _lib.check(_lib.lib().sdfr_backward(
            grad_depth_image.data_ptr(), depth_image.data_ptr(), sdf.data_ptr(), int(sdf.shape[-1]), 0,
            _lib.LAYOUT_DENSE, position.data_ptr(), orientation.data_ptr(), inv_scale.data_ptr(), 1, int(width),
            int(height), float(cx), float(cy), float(fx), float(fy), g_sdf.data_ptr(), 0, g_p.data_ptr(),
            g_q.data_ptr(), g_is.data_ptr(), _lib.GRAD_ALL | _lib.ZERO_GRADS, None, _stream()), "sdfr_backward")
    return [g_sdf, g_p, g_q, g_is]
