"""PyTorch interface of the B200 differentiable SDF depth renderer.

Drop-in for ``sdfest.differentiable_renderer`` (reference ``sdf_renderer.py``): the same
``Camera``, ``render_depth_gpu`` and ``SDFRendererFunctionGPU`` names, argument meaning and error
behaviour, so that ``sdfest.estimation``'s render-and-compare loop and the VAE decoder output plug
in unchanged -- but every call lands in ``libsdfrender.so`` (hand-written sm_100a kernels behind a
C ABI, ``include/sdfrender.h``) instead of the reference's pybind/ATen extension.

On top of the reference API this module adds what the reference lacks and the B200 needs to be
kept busy: batched rendering over hypotheses, a fused render-and-compare operator and a
multi-object composite.  There is no CPU path here (the reference's numpy ``render_depth`` lives
on only as the test oracle under ``oracle/``).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from .. import _lib

_SDF_GRAD_MODES = ("reference", "exact")
_default_sdf_grad_mode = "reference"


def set_sdf_grad_mode(mode: str) -> None:
    """Select the corner weights used for SDF gradients by default.

    ``"reference"`` reproduces the reference CUDA kernel (sdf_renderer_cuda.cu:373-388, a
    permutation of the trilinear weights); ``"exact"`` uses the true trilinear weights of the
    reference CPU renderer (simple_renderer.py:399-408).
    """
    global _default_sdf_grad_mode
    if mode not in _SDF_GRAD_MODES:
        raise ValueError(f"sdf_grad_mode must be one of {_SDF_GRAD_MODES}")
    _default_sdf_grad_mode = mode


def get_sdf_grad_mode() -> str:
    return _default_sdf_grad_mode


_SDF_LAYOUT_POLICIES = ("auto", "dense", "skewed")
_sdf_layout_policy = "auto"


def set_sdf_layout_policy(policy: str) -> None:
    """How the kernels read the SDF grids.

    ``"dense"``: straight from the caller's ``(R,R,R)`` tensor (the reference layout).
    ``"skewed"``: from a pitched scratch copy made by one streaming pass (``sdfr_skew_grids``) in
    which voxels a few cells apart fall into different L1 banks -- the 8-corner gathers of a warp
    then cost ~2 L1 cycles instead of ~6; same fp32 values, identical
    results.  ``"auto"`` (default): skewed when the rendering work outweighs the copy.
    """
    global _sdf_layout_policy
    if policy not in _SDF_LAYOUT_POLICIES:
        raise ValueError(f"sdf layout policy must be one of {_SDF_LAYOUT_POLICIES}")
    _sdf_layout_policy = policy


def get_sdf_layout_policy() -> str:
    return _sdf_layout_policy


_skew_geometry = {}


def _skewed_elems(R: int) -> int:
    if R not in _skew_geometry:
        import ctypes

        n = ctypes.c_longlong(0)
        _lib.check(_lib.lib().sdfr_skewed_pitches(R, None, None, ctypes.byref(n)),
                   "sdfr_skewed_pitches")
        _skew_geometry[R] = int(n.value)
    return _skew_geometry[R]


def _grid_operand(sdf: torch.Tensor, R: int, stride: int, n_render: int, pixels: int):
    """(tensor to read, its per-hypothesis stride, layout id) for a render of ``n_render`` images
    of ``pixels`` pixels from ``sdf``; makes the skewed copy when the policy says so."""
    n_grids = 1 if stride == 0 else n_render
    policy = _sdf_layout_policy
    if policy == "dense" or (policy == "auto" and n_render * pixels < n_grids * R ** 3):
        return sdf, stride, _lib.LAYOUT_DENSE
    elems = _skewed_elems(R)
    skewed = torch.empty((n_grids, elems), dtype=torch.float32, device=sdf.device)
    _lib.check(_lib.lib().sdfr_skew_grids(sdf.data_ptr(), R, stride, n_grids, skewed.data_ptr(),
                                          elems, _stream()), "sdfr_skew_grids")
    return skewed, (0 if stride == 0 else elems), _lib.LAYOUT_SKEWED


_EMPTY_SPACE_POLICIES = ("auto", "on", "off")
_empty_space_policy = "auto"


def set_empty_space_policy(policy: str) -> None:
    """Whether renders first bound the part of each grid where a hit is possible at all.

    The reference marches every ray that enters the grid's [-1,1]^3 box (sdf_renderer_cuda.cu:
    272-293); most of them cross only space where the field stays above the hit threshold.
    ``sdfr_grid_bounds`` (one read of the grids) finds the box of cells where the march can terminate;
    rays that miss it are written as 0 without marching, every other ray is marched exactly as before --
    identical images and gradients.  ``"auto"`` (default): when the rendering work outweighs the pass;
    ``"on"`` / ``"off"``: always / never.
    """
    global _empty_space_policy
    if policy not in _EMPTY_SPACE_POLICIES:
        raise ValueError(f"empty-space policy must be one of {_EMPTY_SPACE_POLICIES}")
    _empty_space_policy = policy


def get_empty_space_policy() -> str:
    return _empty_space_policy


def grid_bounds(src: torch.Tensor, R: int, src_stride: int, layout: int, position: torch.Tensor,
                inv_scale: torch.Tensor, threshold: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Cell bounds (``sdfr_cell_bounds`` records, int32 ``(n_grids, 8)``) of the grid operand ``src``
    (as ``_grid_operand`` returns it) for a render of ``position`` / ``inv_scale`` at ``threshold``."""
    B = int(inv_scale.numel())
    n_grids = 1 if src_stride == 0 else B
    if out is None:
        out = torch.empty((n_grids, 8), dtype=torch.int32, device=src.device)
    _lib.check(_lib.lib().sdfr_grid_bounds(src.data_ptr(), R, src_stride, layout, position.data_ptr(),
                                           inv_scale.data_ptr(), B, float(threshold), out.data_ptr(),
                                           _stream()), "sdfr_grid_bounds")
    return out


def _bounds_operand(src, R, src_stride, layout, position, inv_scale, threshold, n_render, pixels):
    """Bounds tensor for a render of ``n_render`` images of ``pixels`` pixels, or None (policy)."""
    n_grids = 1 if src_stride == 0 else n_render
    policy = _empty_space_policy
    if policy == "off" or (policy == "auto" and n_render * pixels < 2 * n_grids * R ** 3):
        return None
    return grid_bounds(src, R, src_stride, layout, position, inv_scale, threshold)


class Camera:
    """Pinhole camera parameters (reference sdf_renderer.py:31-133).

    Converts between pixel-centre conventions: a discrete pixel (x, y) corresponds to the
    continuous image coordinate (x + pixel_center, y + pixel_center).
    """

    def __init__(self, width: int, height: int, fx: float, fy: float, cx: float, cy: float,
                 s: float = 0.0, pixel_center: float = 0.0):
        self.fx = fx
        self.fy = fy
        self.cx = cx
        self.cy = cy
        self.pixel_center = pixel_center
        self.s = s
        self.width = width
        self.height = height

    def get_o3d_pinhole_camera_parameters(self):
        """Open3D pinhole parameters (pixel_center 0, no skew); imports open3d lazily."""
        import numpy as np
        import open3d as o3d

        fx, fy, cx, cy, _ = self.get_pinhole_camera_parameters(0)
        params = o3d.camera.PinholeCameraParameters()
        params.intrinsic.set_intrinsics(self.width, self.height, fx, fy, cx, cy)
        params.extrinsic = np.eye(4)
        return params

    def get_pinhole_camera_parameters(self, pixel_center: float) -> Tuple:
        """(fx, fy, cx, cy, s) with the principal point expressed for ``pixel_center``."""
        cx_corrected = self.cx - self.pixel_center + pixel_center
        cy_corrected = self.cy - self.pixel_center + pixel_center
        return self.fx, self.fy, cx_corrected, cy_corrected, self.s


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def _check_input(t: torch.Tensor, name: str, min_numel: int = 0) -> None:
    """CHECK_INPUT of the reference binding (sdf_renderer.cpp:9-13) plus dtype/size checks."""
    try:  # the common case in one expression (this runs five times per forward+backward)
        if t.is_cuda and t.dtype is torch.float32 and t.is_contiguous() and t.numel() >= min_numel:
            return
    except AttributeError:
        pass
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32 (the renderer is fp32-only), got {t.dtype}")
    if t.numel() < min_numel:
        raise RuntimeError(f"{name} needs at least {min_numel} elements, got {t.numel()}")


def _check_grid(sdf: torch.Tensor, batched: bool) -> int:
    nd = 4 if batched else 3
    if sdf.dim() != nd or not (sdf.shape[-1] == sdf.shape[-2] == sdf.shape[-3]):
        raise RuntimeError(
            f"sdf must have shape {'(B,' if batched else '('}R,R,R), got {tuple(sdf.shape)}")
    if sdf.shape[-1] < 2:
        raise RuntimeError("sdf resolution must be >= 2")
    return int(sdf.shape[-1])


def _camera_params(camera: Camera):
    fx, fy, cx, cy, _ = camera.get_pinhole_camera_parameters(0.5)
    return int(camera.width), int(camera.height), float(cx), float(cy), float(fx), float(fy)


try:  # the raw handle of the current stream without building a torch.cuda.Stream object (20 us -> 0.4 us:
    # two of these per forward+backward were a quarter of the wrapper's CPU time, scripts/dbg/api_profile.py)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover  (a torch build without the private accessor)
    _raw_stream = None


def _stream() -> int:
    """cudaStream_t of the current stream of the current device, as an integer for the C ABI."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _NoCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_CTX = _NoCtx()


def _on_device_of(t: torch.Tensor):
    """Device guard (OptionalCUDAGuard of sdf_renderer.cpp:58, 82) that costs nothing in the
    common case of the tensor living on the current device."""
    if t.device.index == torch.cuda.current_device():
        return _NO_CTX
    return torch.cuda.device_of(t)


def _grad_flags(needs, mode: Optional[str]) -> int:
    mode = _default_sdf_grad_mode if mode is None else mode
    if mode not in _SDF_GRAD_MODES:
        raise ValueError(f"sdf_grad_mode must be one of {_SDF_GRAD_MODES}")
    flags = _lib.SDF_GRAD_EXACT if mode == "exact" else 0
    for need, bit in zip(needs, (_lib.GRAD_SDF, _lib.GRAD_POSITION, _lib.GRAD_ORIENTATION,
                                 _lib.GRAD_INV_SCALE)):
        if need:
            flags |= bit
    return flags


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------------
# reference API: one object, one image
# --------------------------------------------------------------------------------------------
class SDFRendererFunctionGPU(torch.autograd.Function):
    """Renderer function for signed distance fields (reference sdf_renderer.py:267-357)."""

    @staticmethod
    def forward(ctx, sdf, position, orientation, inv_scale, threshold=0.0, camera=None,
                sdf_grad_mode=None):
        """Render the depth image of a 7-DOF discrete signed distance field.

        sdf (R,R,R); position >=3 elements; orientation >=4 elements (x,y,z,w, unit);
        inv_scale >=1 element; any shapes the reference accepts through its raw pointers
        (``(3,)``/``(1,3)``/0-dim ...).  Returns depth (H,W), 0 = no hit.
        """
        if camera is None:
            raise ValueError("camera must be provided")
        _check_input(sdf, "sdf")
        _check_input(position, "position", 3)
        _check_input(orientation, "orientation", 4)
        _check_input(inv_scale, "inv_scale", 1)
        R = _check_grid(sdf, batched=False)
        W, H, cx, cy, fx, fy = _camera_params(camera)
        with _on_device_of(sdf):
            image = torch.empty((H, W), dtype=torch.float32, device=sdf.device)
            # one image: the layout pass would cost about what it saves, unless asked for
            src, layout = sdf, _lib.LAYOUT_DENSE
            if _sdf_layout_policy == "skewed":
                src, _, layout = _grid_operand(sdf, R, 0, 1, W * H)
            _lib.check(_lib.lib().sdfr_forward(
                src.data_ptr(), R, 0, layout, position.data_ptr(), orientation.data_ptr(),
                inv_scale.data_ptr(), 1, W, H, cx, cy, fx, fy, float(threshold),
                image.data_ptr(), None, _stream()), "sdfr_forward")
        ctx.save_for_backward(image, sdf, position, orientation, inv_scale)
        ctx.cam = (W, H, cx, cy, fx, fy)
        ctx.sdf_grad_mode = sdf_grad_mode
        return image

    @staticmethod
    def backward(ctx, grad_depth_image):
        """Gradients w.r.t. sdf, position, orientation, inv_scale (shaped like the inputs).

        Unlike the reference (sdf_renderer.py:346-357) only the gradients autograd asks for are
        computed (``ctx.needs_input_grad``).
        """
        image, sdf, position, orientation, inv_scale = ctx.saved_tensors
        W, H, cx, cy, fx, fy = ctx.cam
        needs = ctx.needs_input_grad[:4]
        flags = _grad_flags(needs, ctx.sdf_grad_mode)
        if position.numel() == 3 and orientation.numel() == 4 and inv_scale.numel() == 1:
            alloc = torch.empty_like  # the library clears exactly these elements itself
            flags |= _lib.ZERO_GRADS
        else:  # over-long pose tensors: the tail must read as zero
            alloc = torch.zeros_like
        g_sdf = alloc(sdf) if needs[0] else None
        g_p = alloc(position) if needs[1] else None
        g_q = alloc(orientation) if needs[2] else None
        g_is = alloc(inv_scale) if needs[3] else None
        if any(needs):
            grad_depth_image = grad_depth_image.contiguous()
            _check_input(grad_depth_image, "grad_depth_image")
            with _on_device_of(sdf):
                _lib.check(_lib.lib().sdfr_backward(
                    grad_depth_image.data_ptr(), image.data_ptr(), sdf.data_ptr(),
                    int(sdf.shape[-1]), 0, _lib.LAYOUT_DENSE, position.data_ptr(),
                    orientation.data_ptr(),
                    inv_scale.data_ptr(), 1, W, H, cx, cy, fx, fy, _ptr(g_sdf), 0, _ptr(g_p),
                    _ptr(g_q), _ptr(g_is), flags, None, _stream()), "sdfr_backward")
        return g_sdf, g_p, g_q, g_is, None, None, None


def render_depth_gpu(sdf: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor,
                     inv_scale: torch.Tensor, width: Optional[int] = None,
                     height: Optional[int] = None, fov_deg: Optional[float] = None,
                     threshold: Optional[float] = 0.0, camera: Optional[Camera] = None, *,
                     sdf_grad_mode: Optional[str] = None):
    """Render depth image of a 7-DOF discrete signed distance field on the GPU.

    Same contract as the reference (sdf_renderer.py:360-424): the SDF pose is in the OpenGL camera
    frame (camera looks along -z, y up, x right); the image follows the computer-vision
    convention (first row is up).  Specify the camera either via ``camera`` or via
    ``width``+``height``+``fov_deg`` (horizontal field of view, square pixels).
    """
    if None not in [width, height, fov_deg] and camera is not None:
        raise ValueError("Either width+height+fov_dev or camera must be provided.")
    if camera is None:
        if None in [width, height, fov_deg]:
            raise ValueError("Either width+height+fov_dev or camera must be provided.")
        f = width / math.tan(fov_deg * math.pi / 180.0 / 2.0) / 2
        camera = Camera(width, height, f, f, width / 2, height / 2, pixel_center=0.5)
    return SDFRendererFunctionGPU.apply(sdf, position, orientation, inv_scale, threshold, camera,
                                        sdf_grad_mode)


def render_depth(*args, **kwargs):
    """The reference's numpy CPU renderer (sdf_renderer.py:136-264) is not part of this package.

    It survives as the test oracle (``oracle/``); the product has no CPU fallback.
    """
    raise NotImplementedError(
        "sdfest_b200 has no CPU renderer; use render_depth_gpu (the numpy path of the reference "
        "is kept only as the test oracle under oracle/)")


# --------------------------------------------------------------------------------------------
# batched extensions
# --------------------------------------------------------------------------------------------
def _check_batch(sdf, position, orientation, inv_scale):
    _check_input(sdf, "sdf")
    _check_input(position, "position")
    _check_input(orientation, "orientation")
    _check_input(inv_scale, "inv_scale")
    if position.dim() != 2 or position.shape[1] != 3:
        raise RuntimeError(f"position must have shape (B,3), got {tuple(position.shape)}")
    B = int(position.shape[0])
    if tuple(orientation.shape) != (B, 4):
        raise RuntimeError(f"orientation must have shape ({B},4), got {tuple(orientation.shape)}")
    if inv_scale.numel() != B:
        raise RuntimeError(f"inv_scale must have {B} elements, got {inv_scale.numel()}")
    if sdf.dim() == 3:
        R = _check_grid(sdf, batched=False)
        stride = 0
    else:
        R = _check_grid(sdf, batched=True)
        if sdf.shape[0] == 1:
            stride = 0
        elif sdf.shape[0] == B:
            stride = R * R * R
        else:
            raise RuntimeError(f"sdf batch must be 1 or {B}, got {sdf.shape[0]}")
    return B, R, stride


class _BatchedRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, position, orientation, inv_scale, threshold, camera, sdf_grad_mode):
        B, R, stride = _check_batch(sdf, position, orientation, inv_scale)
        W, H, cx, cy, fx, fy = _camera_params(camera)
        with _on_device_of(sdf):
            depth = torch.empty((B, H, W), dtype=torch.float32, device=sdf.device)
            src, src_stride, layout = _grid_operand(sdf, R, stride, B, W * H)
            bounds = _bounds_operand(src, R, src_stride, layout, position, inv_scale, threshold, B, W * H)
            _lib.check(_lib.lib().sdfr_forward(
                src.data_ptr(), R, src_stride, layout, position.data_ptr(),
                orientation.data_ptr(), inv_scale.data_ptr(), B, W, H, cx, cy, fx, fy,
                float(threshold), depth.data_ptr(), _ptr(bounds), _stream()), "sdfr_forward")
        ctx.save_for_backward(depth, sdf, position, orientation, inv_scale)
        ctx.meta = (B, R, stride, W, H, cx, cy, fx, fy, sdf_grad_mode)
        ctx.bounds = bounds  # valid for this depth image: the backward scans a smaller region
        return depth

    @staticmethod
    def backward(ctx, grad_depth):
        depth, sdf, position, orientation, inv_scale = ctx.saved_tensors
        B, R, stride, W, H, cx, cy, fx, fy, mode = ctx.meta
        needs = ctx.needs_input_grad[:4]
        flags = _grad_flags(needs, mode)
        g_sdf = torch.empty_like(sdf) if needs[0] else None
        g_p = torch.empty_like(position) if needs[1] else None
        g_q = torch.empty_like(orientation) if needs[2] else None
        g_is = torch.empty_like(inv_scale) if needs[3] else None
        flags |= _lib.ZERO_GRADS
        if any(needs):
            grad_depth = grad_depth.contiguous()
            _check_input(grad_depth, "grad_depth")
            with _on_device_of(sdf):
                _lib.check(_lib.lib().sdfr_backward(
                    grad_depth.data_ptr(), depth.data_ptr(), sdf.data_ptr(), R, stride,
                    _lib.LAYOUT_DENSE, position.data_ptr(), orientation.data_ptr(),
                    inv_scale.data_ptr(), B, W, H, cx, cy, fx, fy, _ptr(g_sdf), stride,
                    _ptr(g_p), _ptr(g_q), _ptr(g_is), flags, _ptr(ctx.bounds), _stream()), "sdfr_backward")
        return g_sdf, g_p, g_q, g_is, None, None, None


def render_depth_batched(sdf: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor,
                         inv_scale: torch.Tensor, threshold: float, camera: Camera, *,
                         sdf_grad_mode: Optional[str] = None) -> torch.Tensor:
    """Render B pose/shape hypotheses in one launch.

    sdf (R,R,R) or (1,R,R,R) shared by all hypotheses, or (B,R,R,R); position (B,3);
    orientation (B,4); inv_scale (B,) -> depth (B,H,W).  Differentiable w.r.t. all four; a
    shared grid receives the sum of the per-hypothesis SDF gradients.
    """
    return _BatchedRender.apply(sdf, position, orientation, inv_scale, threshold, camera,
                                sdf_grad_mode)


class _RenderAndCompare(torch.autograd.Function):
    """Masked-L1 render-and-compare.  When a gradient will be needed, the backward is folded into
    the forward traversal (sdfr_compare_fused: one kernel, unnormalised gradients) and backward()
    only applies the per-hypothesis factor upstream/n_overlap (sdfr_scale_grads)."""

    @staticmethod
    def forward(ctx, sdf, position, orientation, inv_scale, depth_obs, threshold, camera,
                sdf_grad_mode):
        B, R, stride = _check_batch(sdf, position, orientation, inv_scale)
        W, H, cx, cy, fx, fy = _camera_params(camera)
        _check_input(depth_obs, "depth_obs")
        if tuple(depth_obs.shape) == (H, W):
            obs_stride = 0
        elif tuple(depth_obs.shape) == (B, H, W):
            obs_stride = H * W
        else:
            raise RuntimeError(
                f"depth_obs must have shape ({H},{W}) or ({B},{H},{W}), got {tuple(depth_obs.shape)}")
        needs = tuple(ctx.needs_input_grad[:4])
        # a grid shared by several hypotheses cannot take the deferred per-hypothesis scaling
        fused = any(needs) and not (needs[0] and stride == 0 and B > 1)
        grads = [None, None, None, None]
        with _on_device_of(sdf):
            depth = torch.empty((B, H, W), dtype=torch.float32, device=sdf.device)
            sums = torch.empty((2, B), dtype=torch.float32, device=sdf.device)
            lib = _lib.lib()
            src, src_stride, layout = _grid_operand(sdf, R, stride, B, W * H)
            bounds = _bounds_operand(src, R, src_stride, layout, position, inv_scale, threshold, B, W * H)
            if fused:
                flags = _grad_flags(needs, sdf_grad_mode) | _lib.ZERO_GRADS
                grads = [torch.empty_like(t) if n else None
                         for t, n in zip((sdf, position, orientation, inv_scale), needs)]
                _lib.check(lib.sdfr_compare_fused(
                    src.data_ptr(), R, src_stride, layout, position.data_ptr(),
                    orientation.data_ptr(), inv_scale.data_ptr(), B, W, H, cx, cy, fx, fy,
                    float(threshold),
                    depth_obs.data_ptr(), obs_stride, depth.data_ptr(), sums[0].data_ptr(),
                    sums[1].data_ptr(), _ptr(grads[0]), stride, _ptr(grads[1]), _ptr(grads[2]),
                    _ptr(grads[3]), flags, _ptr(bounds), _stream()), "sdfr_compare_fused")
            else:
                _lib.check(lib.sdfr_compare_forward(
                    src.data_ptr(), R, src_stride, layout, position.data_ptr(),
                    orientation.data_ptr(), inv_scale.data_ptr(), B, W, H, cx, cy, fx, fy,
                    float(threshold),
                    depth_obs.data_ptr(), obs_stride, depth.data_ptr(), sums[0].data_ptr(),
                    sums[1].data_ptr(), _lib.ZERO_GRADS, _ptr(bounds), _stream()), "sdfr_compare_forward")
            loss = sums[0] / sums[1]  # NaN where nothing overlaps, as torch.mean of an empty set
            n_overlap = sums[1].clone()
        ctx.save_for_backward(depth, depth_obs, sums, sdf, position, orientation, inv_scale)
        ctx.meta = (B, R, stride, obs_stride, W, H, cx, cy, fx, fy, sdf_grad_mode)
        ctx.fused_grads = grads if fused else None
        ctx.bounds = bounds
        ctx.mark_non_differentiable(depth, n_overlap)
        return loss, depth, n_overlap

    @staticmethod
    def backward(ctx, grad_loss, _grad_depth, _grad_n):
        depth, depth_obs, sums, sdf, position, orientation, inv_scale = ctx.saved_tensors
        B, R, stride, obs_stride, W, H, cx, cy, fx, fy, mode = ctx.meta
        needs = ctx.needs_input_grad[:4]
        if not any(needs):
            return (None,) * 8
        flags = _grad_flags(needs, mode)
        upstream = grad_loss.to(torch.float32).contiguous()
        lib = _lib.lib()
        with _on_device_of(sdf):
            if ctx.fused_grads is not None:
                # the traversal already happened: only the deferred upstream/n_overlap factor is left
                g_sdf, g_p, g_q, g_is = ctx.fused_grads
                ctx.fused_grads = None  # scaled in place: a second backward re-traverses below
                _lib.check(lib.sdfr_scale_grads(
                    sums[1].data_ptr(), upstream.data_ptr(), R, B, _ptr(g_sdf), stride, _ptr(g_p),
                    _ptr(g_q), _ptr(g_is), flags, None, 0, _stream()), "sdfr_scale_grads")
            else:
                g_sdf = torch.empty_like(sdf) if needs[0] else None
                g_p = torch.empty_like(position) if needs[1] else None
                g_q = torch.empty_like(orientation) if needs[2] else None
                g_is = torch.empty_like(inv_scale) if needs[3] else None
                _lib.check(lib.sdfr_compare_backward(
                    depth.data_ptr(), depth_obs.data_ptr(), obs_stride, sums[1].data_ptr(),
                    upstream.data_ptr(), sdf.data_ptr(), R, stride, _lib.LAYOUT_DENSE,
                    position.data_ptr(), orientation.data_ptr(), inv_scale.data_ptr(), B, W, H,
                    cx, cy, fx, fy,
                    _ptr(g_sdf), stride, _ptr(g_p), _ptr(g_q), _ptr(g_is),
                    flags | _lib.ZERO_GRADS, _ptr(ctx.bounds), _stream()), "sdfr_compare_backward")
        return g_sdf, g_p, g_q, g_is, None, None, None, None


def render_and_compare(sdf: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor,
                       inv_scale: torch.Tensor, depth_obs: torch.Tensor, threshold: float,
                       camera: Camera, *, sdf_grad_mode: Optional[str] = None):
    """Fused render + masked-L1 depth loss of the reference pipeline, batched.

    For every hypothesis b:  ``loss[b] = mean |depth[b] - depth_obs|`` over pixels where both are
    positive (estimation/simple_setup.py:125-131).  Returns ``(loss (B,), depth (B,H,W),
    n_overlap (B,))``; only ``loss`` is differentiable, and its backward never materialises a
    grad_depth image.  ``depth_obs`` is (H,W) (shared) or (B,H,W).
    """
    return _RenderAndCompare.apply(sdf, position, orientation, inv_scale, depth_obs, threshold,
                                   camera, sdf_grad_mode)


class _CompositeRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, position, orientation, inv_scale, threshold, camera, sdf_grad_mode):
        K, R, stride = _check_batch(sdf, position, orientation, inv_scale)
        W, H, cx, cy, fx, fy = _camera_params(camera)
        with _on_device_of(sdf):
            depth = torch.empty((H, W), dtype=torch.float32, device=sdf.device)
            winner = torch.empty((H, W), dtype=torch.int32, device=sdf.device)
            src, src_stride, layout = _grid_operand(sdf, R, stride, K, W * H // max(K, 1))
            bounds = _bounds_operand(src, R, src_stride, layout, position, inv_scale, threshold, K,
                                     W * H // max(K, 1))
            _lib.check(_lib.lib().sdfr_forward_composite(
                src.data_ptr(), R, src_stride, layout, position.data_ptr(),
                orientation.data_ptr(), inv_scale.data_ptr(), K, W, H, cx, cy, fx, fy,
                float(threshold),
                depth.data_ptr(), winner.data_ptr(), _ptr(bounds), _stream()), "sdfr_forward_composite")
        ctx.save_for_backward(depth, winner, sdf, position, orientation, inv_scale)
        ctx.meta = (K, R, stride, W, H, cx, cy, fx, fy, sdf_grad_mode)
        ctx.mark_non_differentiable(winner)
        return depth, winner

    @staticmethod
    def backward(ctx, grad_depth, _grad_winner):
        depth, winner, sdf, position, orientation, inv_scale = ctx.saved_tensors
        K, R, stride, W, H, cx, cy, fx, fy, mode = ctx.meta
        needs = ctx.needs_input_grad[:4]
        flags = _grad_flags(needs, mode)
        g_sdf = torch.empty_like(sdf) if needs[0] else None
        g_p = torch.empty_like(position) if needs[1] else None
        g_q = torch.empty_like(orientation) if needs[2] else None
        g_is = torch.empty_like(inv_scale) if needs[3] else None
        flags |= _lib.ZERO_GRADS
        if any(needs):
            grad_depth = grad_depth.contiguous()
            _check_input(grad_depth, "grad_depth")
            with _on_device_of(sdf):
                _lib.check(_lib.lib().sdfr_backward_composite(
                    grad_depth.data_ptr(), depth.data_ptr(), winner.data_ptr(), sdf.data_ptr(), R,
                    stride, _lib.LAYOUT_DENSE, position.data_ptr(), orientation.data_ptr(),
                    inv_scale.data_ptr(), K,
                    W, H, cx, cy, fx, fy, _ptr(g_sdf), stride, _ptr(g_p), _ptr(g_q), _ptr(g_is),
                    flags, _stream()), "sdfr_backward_composite")
        return g_sdf, g_p, g_q, g_is, None, None, None


def render_depth_composite(sdf: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor,
                           inv_scale: torch.Tensor, threshold: float, camera: Camera, *,
                           sdf_grad_mode: Optional[str] = None):
    """Render K posed objects into ONE depth map (per-pixel minimum positive depth).

    Returns ``(depth (H,W), winner (H,W) int32)``; ``winner`` is the index of the object seen at
    each pixel (-1 = background).  Gradients flow to the winning object of each pixel.
    """
    return _CompositeRender.apply(sdf, position, orientation, inv_scale, threshold, camera,
                                  sdf_grad_mode)


def forward_stats(sdf, position, orientation, inv_scale, threshold, camera, empty_space: bool = False):
    """Work counters of a batched render: dict(samples, box_pixels, hit_pixels, capped_rays).

    ``samples`` is the S and ``hit_pixels`` the Hh of the roofline's algorithmic-bytes formula
    (DESIGN.md); the depth image is rendered as a side effect and discarded.  ``box_pixels`` counts the
    rays that are marched: all rays entering the grid's box, or -- ``empty_space=True`` -- only those
    that enter the empty-space bounds of ``sdfr_grid_bounds``.
    """
    B, R, stride = _check_batch(sdf, position, orientation, inv_scale)
    W, H, cx, cy, fx, fy = _camera_params(camera)
    with _on_device_of(sdf):
        depth = torch.empty((B, H, W), dtype=torch.float32, device=sdf.device)
        stats = torch.zeros(4, dtype=torch.int64, device=sdf.device)
        bounds = grid_bounds(sdf, R, stride, _lib.LAYOUT_DENSE, position, inv_scale, threshold) \
            if empty_space else None
        _lib.check(_lib.lib().sdfr_forward_stats(
            sdf.data_ptr(), R, stride, _lib.LAYOUT_DENSE, position.data_ptr(), orientation.data_ptr(),
            inv_scale.data_ptr(), B, W, H, cx, cy, fx, fy, float(threshold), depth.data_ptr(),
            stats.data_ptr(), _ptr(bounds), _stream()), "sdfr_forward_stats")
        s = stats.tolist()
    return {"samples": s[0], "box_pixels": s[1], "hit_pixels": s[2], "capped_rays": s[3]}
