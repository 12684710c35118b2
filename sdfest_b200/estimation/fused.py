"""One loop iteration's device work between the decoder trunk and the optimiser as FIVE kernels.

The reference evaluates, per hypothesis and per iteration (estimation/simple_setup.py:408-456):
decode (…, interpolate, conv 1x1; vae/sdf_vae.py:235-247) -> render (sdf_renderer_cuda.cu) ->
masked L1 (simple_setup.py:125-131) -> pc_loss (estimation/losses.py:32-135) -> autograd backward
through all of them.  Composed from this package's separate operators that is already batched and
sync-free, but still pays for layout glue between them: a dense grid that is re-copied into the
skewed layout, a pass that scales the raw SDF gradient by upstream/n_overlap, a second SDF-gradient
grid from the point loss and the elementwise add of the two.  ``decode_render_compare`` chains the
C-ABI kernels directly instead:

  forward   sdfr_decoder_tail_forward   x (B,C,S^3) -> grids, written straight in the skewed layout
            sdfr_compare_fused          render + masked L1 + its raw (unnormalised) backward
            sdfr_point_loss_forward     reads the same skewed grids
  backward  sdfr_point_loss_backward    its SDF gradient goes to a second dense grid
            sdfr_decoder_tail_backward  g_x = W^T (coef * g_render + g_points): the deferred
                                        normalisation and the sum of the two grids ride on its load
            sdfr_scale_grads            on the 8 pose gradients per hypothesis only

so the R^3 grids and gradient grids are each written once and read once.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib
from ..differentiable_renderer.sdf_renderer import (_camera_params, _check_input, _grad_flags,
                                                    _on_device_of, _ptr, _skewed_elems, _stream)


class _DecodeRenderCompare(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, base, position, orientation, scale, depth_obs, points,
                resolution, threshold, camera, depth_weight, pc_weight, sdf_grad_mode):
        for t, n in ((x, "x"), (weight, "weight"), (position, "position"),
                     (orientation, "orientation"), (scale, "scale"), (depth_obs, "depth_obs")):
            _check_input(t, n)
        if x.dim() != 5 or not (x.shape[2] == x.shape[3] == x.shape[4]):
            raise RuntimeError(f"x must have shape (B,C,S,S,S), got {tuple(x.shape)}")
        B, C, S = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        R = int(resolution)
        if weight.numel() != C:
            raise RuntimeError(f"weight must have {C} elements, got {weight.numel()}")
        if tuple(position.shape) != (B, 3) or tuple(orientation.shape) != (B, 4) or scale.numel() != B:
            raise RuntimeError("position (B,3), orientation (B,4), scale (B,) expected")
        W, H, cx, cy, fx, fy = _camera_params(camera)
        if tuple(depth_obs.shape) == (H, W):
            obs_stride = 0
        elif tuple(depth_obs.shape) == (B, H, W):
            obs_stride = H * W
        else:
            raise RuntimeError(f"depth_obs must have shape ({H},{W}) or ({B},{H},{W})")
        if base is not None:
            _check_input(base, "base", R ** 3)
        M = 0
        if points is not None and pc_weight:
            _check_input(points, "points")
            if points.dim() != 2 or points.shape[1] != 3:
                raise RuntimeError(f"points must have shape (M,3), got {tuple(points.shape)}")
            M = int(points.shape[0])
        if weight.requires_grad or (bias is not None and bias.requires_grad):
            raise RuntimeError("decode_render_compare is for a frozen decoder tail")
        needs = ctx.needs_input_grad
        need_x = needs[0]
        need_pose = (needs[4], needs[5], needs[6])
        lib = _lib.lib()
        with _on_device_of(x):
            dev = x.device
            SK = _skewed_elems(R)
            grids = torch.empty((B, SK), dtype=torch.float32, device=dev)
            st = _stream()
            inv_scale = (1.0 / scale.detach()).contiguous()
            from ..differentiable_renderer.sdf_renderer import get_empty_space_policy
            bounds = None
            if get_empty_space_policy() != "off":
                # the tail compares every value with the hypothesis' hit-threshold bound as it stores it:
                # rays that cannot hit anything are not marched (identical images)
                bounds_t = torch.empty((B, 8), dtype=torch.int32, device=dev)
                _lib.check(lib.sdfr_decoder_tail_forward_bounds(
                    x.data_ptr(), C, S, weight.data_ptr(), _ptr(bias), _ptr(base), B, R, grids.data_ptr(), SK,
                    _lib.LAYOUT_SKEWED, position.data_ptr(), inv_scale.data_ptr(), float(threshold),
                    bounds_t.data_ptr(), st), "sdfr_decoder_tail_forward_bounds")
                bounds = bounds_t.data_ptr()
            else:
                _lib.check(lib.sdfr_decoder_tail_forward(
                    x.data_ptr(), C, S, weight.data_ptr(), _ptr(bias), _ptr(base), B, R,
                    grids.data_ptr(), SK, _lib.LAYOUT_SKEWED, st), "sdfr_decoder_tail_forward")
            depth = torch.empty((B, H, W), dtype=torch.float32, device=dev)
            sums = torch.empty((2, B), dtype=torch.float32, device=dev)
            g_sdf = torch.empty((B, R ** 3), dtype=torch.float32, device=dev) if need_x else None
            g_p = torch.empty((B, 3), dtype=torch.float32, device=dev) if need_pose[0] else None
            g_q = torch.empty((B, 4), dtype=torch.float32, device=dev) if need_pose[1] else None
            g_is = torch.empty((B,), dtype=torch.float32, device=dev) if need_pose[2] else None
            rflags = _grad_flags((need_x, *need_pose), sdf_grad_mode)
            if rflags & _lib.GRAD_ALL:
                _lib.check(lib.sdfr_compare_fused(
                    grids.data_ptr(), R, SK, _lib.LAYOUT_SKEWED, position.data_ptr(),
                    orientation.data_ptr(), inv_scale.data_ptr(), B, W, H, cx, cy, fx, fy,
                    float(threshold), depth_obs.data_ptr(), obs_stride, depth.data_ptr(),
                    sums[0].data_ptr(), sums[1].data_ptr(), _ptr(g_sdf), R ** 3, _ptr(g_p), _ptr(g_q),
                    _ptr(g_is), rflags | _lib.ZERO_GRADS, bounds, st), "sdfr_compare_fused")
            else:
                _lib.check(lib.sdfr_compare_forward(
                    grids.data_ptr(), R, SK, _lib.LAYOUT_SKEWED, position.data_ptr(),
                    orientation.data_ptr(), inv_scale.data_ptr(), B, W, H, cx, cy, fx, fy,
                    float(threshold), depth_obs.data_ptr(), obs_stride, depth.data_ptr(),
                    sums[0].data_ptr(), sums[1].data_ptr(), _lib.ZERO_GRADS, bounds, st), "sdfr_compare_forward")
            loss_depth = sums[0] / sums[1]  # NaN where nothing overlaps (torch.mean of an empty set)
            loss = float(depth_weight) * torch.nan_to_num(loss_depth, nan=0.0)
            loss_pc = None
            if M > 0:
                pl = torch.empty((B,), dtype=torch.float32, device=dev)
                _lib.check(lib.sdfr_point_loss_forward(
                    points.data_ptr(), 0, M, grids.data_ptr(), R, SK, _lib.LAYOUT_SKEWED,
                    position.data_ptr(), orientation.data_ptr(), scale.data_ptr(), B, pl.data_ptr(),
                    _lib.ZERO_GRADS, st), "sdfr_point_loss_forward")
                loss_pc = pl / M
                loss = loss + float(pc_weight) * loss_pc
            n_overlap = sums[1].clone()
        ctx.save_for_backward(weight, position, orientation, scale, points if M > 0 else None,
                              grids, sums, g_sdf, g_p, g_q, g_is)
        ctx.meta = (B, C, S, R, M, float(depth_weight), float(pc_weight), need_x, need_pose)
        ctx.mark_non_differentiable(depth, n_overlap, loss_depth)
        return loss, depth, n_overlap, loss_depth

    @staticmethod
    def backward(ctx, grad_loss, _gd, _gn, _gl):
        (weight, position, orientation, scale, points, grids, sums, g_sdf, g_p, g_q,
         g_is) = ctx.saved_tensors
        B, C, S, R, M, dw, pw, need_x, need_pose = ctx.meta
        lib = _lib.lib()
        out = [None] * 15
        if not (need_x or any(need_pose)):
            return tuple(out)
        with _on_device_of(grids):
            dev = grids.device
            st = _stream()
            SK = grids.shape[1]
            grad_loss = grad_loss.to(torch.float32)
            up_d = (grad_loss * dw).contiguous()
            gp_pc = gq_pc = gs_pc = g_sdf_pc = None
            if M > 0:
                up_p = (grad_loss * (pw / M)).contiguous()
                g_sdf_pc = torch.empty((B, R ** 3), dtype=torch.float32, device=dev) if need_x else None
                gp_pc = torch.empty((B, 3), dtype=torch.float32, device=dev) if need_pose[0] else None
                gq_pc = torch.empty((B, 4), dtype=torch.float32, device=dev) if need_pose[1] else None
                gs_pc = torch.empty((B,), dtype=torch.float32, device=dev) if need_pose[2] else None
                flags = _lib.ZERO_GRADS
                for need, bit in zip((need_x, *need_pose), (_lib.GRAD_SDF, _lib.GRAD_POSITION,
                                                            _lib.GRAD_ORIENTATION, _lib.GRAD_INV_SCALE)):
                    if need:
                        flags |= bit
                _lib.check(lib.sdfr_point_loss_backward(
                    points.data_ptr(), 0, M, grids.data_ptr(), R, SK, _lib.LAYOUT_SKEWED,
                    position.data_ptr(), orientation.data_ptr(), scale.data_ptr(), B,
                    up_p.data_ptr(), _ptr(g_sdf_pc), R ** 3, _ptr(gp_pc), _ptr(gq_pc), _ptr(gs_pc),
                    flags, st), "sdfr_point_loss_backward")
            if need_x:
                g_x = torch.empty((B, C, S, S, S), dtype=torch.float32, device=dev)
                _lib.check(lib.sdfr_decoder_tail_backward(
                    g_sdf.data_ptr(), R ** 3, sums[1].data_ptr(), up_d.data_ptr(), _ptr(g_sdf_pc),
                    R ** 3, weight.data_ptr(), C, S, B, R, g_x.data_ptr(), st),
                    "sdfr_decoder_tail_backward")
                out[0] = g_x
            if any(need_pose):
                # the render's pose gradients are still raw: apply upstream/n_overlap out of place
                # (a second backward through the same graph must see the raw values again)
                coef = torch.where(sums[1] > 0, up_d / sums[1], torch.zeros_like(up_d))
                if need_pose[0]:
                    out[4] = g_p * coef[:, None] + (gp_pc if gp_pc is not None else 0.0)
                if need_pose[1]:
                    out[5] = g_q * coef[:, None] + (gq_pc if gq_pc is not None else 0.0)
                if need_pose[2]:  # render differentiates w.r.t. inv_scale = 1/scale
                    gs = g_is * coef * (-1.0 / (scale * scale))
                    out[6] = (gs + gs_pc if gs_pc is not None else gs).view_as(scale)
        return tuple(out)


def decode_render_compare(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                          position: torch.Tensor, orientation: torch.Tensor, scale: torch.Tensor,
                          depth_obs: torch.Tensor, points: Optional[torch.Tensor], resolution: int,
                          threshold: float, camera, *, base: Optional[torch.Tensor] = None,
                          depth_weight: float = 1.0, pc_weight: float = 3.0,
                          sdf_grad_mode: Optional[str] = None):
    """Decoder tail + render-and-compare + point-cloud loss of one loop iteration, batched.

    x (B,C,S,S,S): output of the decoder trunk (``FusedTailDecoder.trunk``); weight (C,), bias
    (1,) or None: the decoder's last 1x1x1 convolution; position (B,3); orientation (B,4) UNIT
    quaternions; scale (B,) (not inverted); depth_obs (H,W) or (B,H,W); points (M,3) observed
    points in the camera frame or None.  Returns ``(loss (B,), depth (B,H,W), n_overlap (B,),
    loss_depth (B,))`` with ``loss = depth_weight * nan_to_num(masked-L1 depth loss) + pc_weight *
    mean |SDF(points) * scale|`` (simple_setup.py:125-144; default weights
    estimation/configs/default.yaml:14-16).  Only ``loss`` is differentiable -- w.r.t. x,
    position, orientation and scale.
    """
    return _DecodeRenderCompare.apply(
        x.contiguous(), weight.reshape(-1).contiguous(), bias, base, position.contiguous(),
        orientation.contiguous(), scale.contiguous(), depth_obs,
        None if points is None else points.contiguous(), resolution, threshold, camera,
        depth_weight, pc_weight, sdf_grad_mode)
