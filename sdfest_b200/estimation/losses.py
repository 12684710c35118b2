"""Batched, sync-free restatements of the reference's caller-side loss helpers.

Reference: estimation/losses.py:32-135 (pc_loss), initialization/pointset_utils.py:34-89
(depth_to_pointcloud), initialization/quaternion_utils.py (scalar-last quaternions).
These stay on PyTorch: they are not the hot path (SURVEY.md section 8f ranks a fused kernel
for them as the next step after the renderer).
"""
from __future__ import annotations

import torch


def depth_to_pointcloud(depth_image: torch.Tensor, camera, mask=None) -> torch.Tensor:
    """Observed depth (H,W) -> points (N,3) in the OpenGL camera frame (x right, y up, z back).

    Uses the pixel-centre-0 principal point like the reference (pointset_utils.py:57).  The
    nonzero() makes this a host sync; it runs once per observation, outside the loop.
    """
    fx, fy, cx, cy, _ = camera.get_pinhole_camera_parameters(0.0)
    img = depth_image if mask is None else depth_image * mask
    rows, cols = torch.nonzero(img, as_tuple=True)
    z = depth_image[rows, cols]
    return torch.stack([(cols.float() - cx) * z / fx, -(rows.float() - cy) * z / fy, -z], dim=1)


def subsample_points(points: torch.Tensor, max_points: int) -> torch.Tensor:
    """At most ``max_points`` of the (N,3) points, a fixed random subset (seed 0); 0 = keep all."""
    if max_points and points.shape[0] > max_points:
        sel = torch.randperm(points.shape[0], device=points.device,
                             generator=torch.Generator(points.device).manual_seed(0))[:max_points]
        points = points[sel]
    return points.contiguous()


PAD_COORDINATE = 1.0e6  # metres: outside every SDF volume, contributes exactly 0 (losses.py:84-101)


def depth_to_pointclouds(depth_images: torch.Tensor, camera, max_points: int = 0):
    """Observed depth maps (K,H,W) of K object instances -> (points (K,M,3), counts (K,)).

    Instance k owns the first counts[k] points of row k (``depth_to_pointcloud`` + ``subsample_points``
    of its own map); the rest of the row is padding far outside any SDF volume, where the reference's
    ``pc_loss`` is exactly 0 (estimation/losses.py:84-101), so a mean over the instance's own points is
    ``sum / counts[k]``.  M = the largest count (at least 1).  One host sync per instance, once per
    observation, outside the loop.
    """
    if depth_images.dim() != 3:
        raise RuntimeError(f"depth_images must be (K,H,W), got {tuple(depth_images.shape)}")
    clouds = [subsample_points(depth_to_pointcloud(d, camera), max_points) for d in depth_images]
    counts = torch.tensor([c.shape[0] for c in clouds], dtype=torch.int64)
    M = max(int(counts.max()) if len(clouds) else 0, 1)
    out = torch.full((len(clouds), M, 3), PAD_COORDINATE, dtype=torch.float32, device=depth_images.device)
    for k, c in enumerate(clouds):
        out[k, : c.shape[0]] = c
    return out, counts.to(depth_images.device)


# --------------------------------------------------------------------------------------------
# CUDA path: the same loss as ONE kernel pair (sdfr_point_loss_forward / _backward)
# --------------------------------------------------------------------------------------------
class _PointLoss(torch.autograd.Function):
    """mean_m |pc_loss(points, ...)[b, m]| per hypothesis, on the GPU through the C ABI."""

    @staticmethod
    def forward(ctx, points, position, orientation, scale, sdf):
        from .. import _lib
        from ..differentiable_renderer.sdf_renderer import _check_input, _on_device_of, _stream

        for t, n in ((points, "points"), (position, "position"), (orientation, "orientation"),
                     (scale, "scale"), (sdf, "sdf")):
            _check_input(t, n)
        B = position.shape[0]
        if position.shape != (B, 3) or orientation.shape != (B, 4) or scale.numel() != B:
            raise RuntimeError("position (B,3), orientation (B,4), scale (B,) expected")
        if points.dim() == 2 and points.shape[1] == 3:
            M, pstride = points.shape[0], 0
        elif points.dim() == 3 and points.shape[0] == B and points.shape[2] == 3:
            M, pstride = points.shape[1], points.shape[1] * 3
        else:
            raise RuntimeError(f"points must be (M,3) or ({B},M,3), got {tuple(points.shape)}")
        R = sdf.shape[-1]
        if sdf.dim() != 4 or sdf.shape[0] not in (1, B) or not (sdf.shape[1] == sdf.shape[2] == R):
            raise RuntimeError(f"sdf must be (1|{B},R,R,R), got {tuple(sdf.shape)}")
        stride = 0 if sdf.shape[0] == 1 else R ** 3
        with _on_device_of(sdf):
            loss_sum = torch.empty(B, dtype=torch.float32, device=sdf.device)
            _lib.check(_lib.lib().sdfr_point_loss_forward(
                points.data_ptr(), pstride, M, sdf.data_ptr(), R, stride, _lib.LAYOUT_DENSE,
                position.data_ptr(), orientation.data_ptr(), scale.data_ptr(), B,
                loss_sum.data_ptr(), _lib.ZERO_GRADS, _stream()), "sdfr_point_loss_forward")
        ctx.save_for_backward(points, position, orientation, scale, sdf)
        ctx.meta = (B, M, pstride, R, stride)
        return loss_sum / max(M, 1)

    @staticmethod
    def backward(ctx, grad_loss):
        from .. import _lib
        from ..differentiable_renderer.sdf_renderer import _on_device_of, _ptr, _stream

        points, position, orientation, scale, sdf = ctx.saved_tensors
        B, M, pstride, R, stride = ctx.meta
        needs = ctx.needs_input_grad
        if needs[0]:
            raise RuntimeError("the point loss is not differentiable w.r.t. the observed points")
        flags = _lib.ZERO_GRADS
        for need, bit in zip(needs[1:], (_lib.GRAD_POSITION, _lib.GRAD_ORIENTATION,
                                         _lib.GRAD_INV_SCALE, _lib.GRAD_SDF)):
            if need:
                flags |= bit
        g_p = torch.empty_like(position) if needs[1] else None
        g_q = torch.empty_like(orientation) if needs[2] else None
        g_s = torch.empty_like(scale) if needs[3] else None
        g_sdf = torch.empty_like(sdf) if needs[4] else None
        if any(needs[1:]):
            upstream = (grad_loss.to(torch.float32) / max(M, 1)).contiguous()
            with _on_device_of(sdf):
                _lib.check(_lib.lib().sdfr_point_loss_backward(
                    points.data_ptr(), pstride, M, sdf.data_ptr(), R, stride, _lib.LAYOUT_DENSE,
                    position.data_ptr(), orientation.data_ptr(), scale.data_ptr(), B,
                    upstream.data_ptr(), _ptr(g_sdf), stride, _ptr(g_p), _ptr(g_q), _ptr(g_s), flags,
                    _stream()), "sdfr_point_loss_backward")
        return None, g_p, g_q, g_s, g_sdf


def point_loss(points: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor,
               scale: torch.Tensor, sdf: torch.Tensor) -> torch.Tensor:
    """Per-hypothesis ``mean_m |pc_loss[b, m]|`` (estimation/simple_setup.py:134-144) through the fused
    kernels of ``libsdfrender.so`` (all observed points, no sub-sampling, no intermediate (B,M)
    tensors).  points (M,3) or (B,M,3); sdf (B|1,R,R,R).  CUDA tensors only: there is no CPU path in
    the product (the torch restatement of the reference's ``pc_loss`` lives in ``oracle/pc_loss.py`` as
    test infrastructure)."""
    if not points.is_cuda:
        raise RuntimeError("point_loss needs CUDA tensors: sdfest_b200 has no CPU fallback")
    return _PointLoss.apply(points.contiguous(), position.contiguous(), orientation.contiguous(),
                            scale.contiguous(), sdf.contiguous())
