"""Latent -> SDF grid decoder with the architecture of the reference's SDFVAE decoder.

NOT a kernel target (BASELINE north star: "the VAE and initialization networks stay on
PyTorch/cuDNN because they are not the hot path").  It exists so that the batched loop and its
benchmark can exercise BASELINE config 2 ("50 Adam steps on pose/scale/latent") with a decoder of
the reference's shape and cost where ``sdfest.vae`` is not importable: fully-connected stack with
ReLU, reshape to (C, s, s, s), then per stage [trilinear interpolate to the stage's input size,
Conv3d, optional ReLU], and a final interpolate to the grid resolution (reference
vae/sdf_vae.py:165-259; default layer sizes = estimation/configs/models/mug.yaml:2-12).
Randomly initialised weights give a near-constant field without an iso-surface, so
``SurfaceDecoder`` adds the decoder output (mean removed) to a fixed analytic shape.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from .. import _lib

MUG_FC = (20, 50, 8192)
MUG_CONV = ((8, 16, 16, 3, True), (16, 16, 8, 3, True), (32, 8, 4, 3, True), (64, 4, 1, 1, False))


class SDFDecoder(nn.Module):
    """conv stages: (in_size, in_channels, out_channels, kernel_size, relu)."""

    def __init__(self, volume_size: int = 64, latent_size: int = 8, fc: Sequence[int] = MUG_FC,
                 conv: Sequence[tuple] = MUG_CONV):
        super().__init__()
        assert fc[-1] == conv[0][1] * conv[0][0] ** 3 and conv[-1][2] == 1
        self.volume_size, self.conv_info = volume_size, tuple(conv)
        sizes = [latent_size, *fc]
        self.fc = nn.ModuleList(nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))
        self.conv = nn.ModuleList(nn.Conv3d(ci, co, k) for _, ci, co, k, _ in conv)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        out = z
        for layer in self.fc:
            out = torch.relu(layer(out))
        s0, c0 = self.conv_info[0][0], self.conv_info[0][1]
        out = out.view(-1, c0, s0, s0, s0)
        for (size, _, _, _, relu), layer in zip(self.conv_info, self.conv):
            if out.shape[2] != size:
                out = nn.functional.interpolate(out, size=(size,) * 3, mode="trilinear",
                                                align_corners=False)
            out = layer(out)
            if relu:
                out = torch.relu(out)
        if out.shape[2] != self.volume_size:
            out = nn.functional.interpolate(out, size=(self.volume_size,) * 3, mode="trilinear",
                                            align_corners=False)
        return out  # (N, 1, D, D, D)


class SurfaceDecoder(nn.Module):
    """``base + (decoder(z) - mean)``: a decodable family that always has a surface."""

    def __init__(self, base: torch.Tensor, decoder: Optional[SDFDecoder] = None, gain: float = 1.0):
        super().__init__()
        self.decoder = decoder if decoder is not None else SDFDecoder(base.shape[-1])
        self.register_buffer("base", base.clone())
        self.gain = gain

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        d = self.decoder(z)
        return self.base[None, None] + self.gain * (d - d.mean(dim=(2, 3, 4), keepdim=True))


# --------------------------------------------------------------------------------------------
# CUDA path: the decoder tail (final trilinear interpolation + 1x1x1 convolution) as one kernel
# --------------------------------------------------------------------------------------------
class _DecoderTail(torch.autograd.Function):
    """``conv1x1(interpolate(x, (R,R,R), "trilinear", align_corners=False))[:, 0]`` through
    ``sdfr_decoder_tail_forward`` / ``_backward`` (reference sdf_vae.py:235-247)."""

    @staticmethod
    def forward(ctx, x, weight, bias, resolution, base=None):
        from ..differentiable_renderer.sdf_renderer import _check_input, _on_device_of, _stream

        _check_input(x, "x")
        _check_input(weight, "weight")
        if x.dim() != 5 or not (x.shape[2] == x.shape[3] == x.shape[4]):
            raise RuntimeError(f"x must have shape (B,C,S,S,S), got {tuple(x.shape)}")
        B, C, S = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        if weight.numel() != C:
            raise RuntimeError(f"weight must have {C} elements, got {weight.numel()}")
        if bias is not None:
            _check_input(bias, "bias", 1)
        if weight.requires_grad or (bias is not None and bias.requires_grad):
            raise RuntimeError("the fused decoder tail is for a frozen decoder (no weight/bias "
                               "gradients; the estimation loop never trains it, simple_setup.py:65)")
        R = int(resolution)
        if base is not None:
            _check_input(base, "base", R ** 3)
        with _on_device_of(x):
            out = torch.empty((B, R, R, R), dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib().sdfr_decoder_tail_forward(
                x.data_ptr(), C, S, weight.data_ptr(), None if bias is None else bias.data_ptr(),
                None if base is None else base.data_ptr(), B, R, out.data_ptr(), R ** 3, _lib.LAYOUT_DENSE, _stream()),
                "sdfr_decoder_tail_forward")
        ctx.save_for_backward(weight)
        ctx.meta = (B, C, S, R)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from ..differentiable_renderer.sdf_renderer import _on_device_of, _stream

        (weight,) = ctx.saved_tensors
        B, C, S, R = ctx.meta
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        grad_out = grad_out.contiguous()
        with _on_device_of(grad_out):
            g_x = torch.empty((B, C, S, S, S), dtype=torch.float32, device=grad_out.device)
            _lib.check(_lib.lib().sdfr_decoder_tail_backward(
                grad_out.data_ptr(), R ** 3, None, None, None, 0, weight.data_ptr(), C, S, B, R,
                g_x.data_ptr(), _stream()), "sdfr_decoder_tail_backward")
        return g_x, None, None, None, None


def decoder_tail(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                 resolution: int, base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused last stage of the reference decoder: x (B,C,S,S,S) -> SDF grids (B,R,R,R).

    Equals ``F.conv3d(F.interpolate(x, (R,R,R), mode="trilinear", align_corners=False),
    weight.view(1,C,1,1,1), bias)[:, 0]`` to fp32 rounding, without the C x R^3 intermediate;
    ``base`` (R,R,R), if given, is added to every grid (residual decoding around a fixed shape).
    CUDA only; differentiable w.r.t. ``x``.
    """
    return _DecoderTail.apply(x.contiguous(), weight.reshape(-1).contiguous(), bias, resolution,
                              None if base is None else base.contiguous())


class _TrunkStage(torch.autograd.Function):
    """One decoder stage (sdf_vae.py:225-247): interpolate to ``in_size`` (if the input is smaller
    or larger), Conv3d 3x3x3 + bias, optional ReLU -- ``sdfr_upsample3d_*`` and ``sdfr_conv3d_*``.
    Differentiable w.r.t. the input only (frozen decoder)."""

    @staticmethod
    def forward(ctx, x, weight, bias, in_size, relu):
        from ..differentiable_renderer.sdf_renderer import _check_input, _on_device_of, _stream

        _check_input(x, "x")
        _check_input(weight, "weight")
        if bias is not None:
            _check_input(bias, "bias")
        if x.dim() != 5 or not (x.shape[2] == x.shape[3] == x.shape[4]):
            raise RuntimeError(f"x must have shape (B,C,S,S,S), got {tuple(x.shape)}")
        if weight.requires_grad or (bias is not None and bias.requires_grad):
            raise RuntimeError("the CUDA decoder trunk is for a frozen decoder (no weight gradients)")
        B, Ci, S = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        Co, k, U = int(weight.shape[0]), int(weight.shape[-1]), int(in_size)
        if tuple(weight.shape) != (Co, Ci, k, k, k):
            raise RuntimeError(f"weight must have shape ({Co},{Ci},k,k,k), got {tuple(weight.shape)}")
        lib = _lib.lib()
        with _on_device_of(x):
            u = x
            if S != U:
                u = torch.empty((B, Ci, U, U, U), dtype=torch.float32, device=x.device)
                _lib.check(lib.sdfr_upsample3d_forward(x.data_ptr(), B * Ci, S, U, u.data_ptr(),
                                                       _stream()), "sdfr_upsample3d_forward")
            O = U - k + 1
            y = torch.empty((B, Co, O, O, O), dtype=torch.float32, device=x.device)
            _lib.check(lib.sdfr_conv3d_forward(
                u.data_ptr(), B, Ci, U, weight.data_ptr(), None if bias is None else bias.data_ptr(),
                Co, k, int(bool(relu)), y.data_ptr(), _stream()), "sdfr_conv3d_forward")
        ctx.save_for_backward(weight, y if relu else None)
        ctx.meta = (B, Ci, S, U, Co, k)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        from ..differentiable_renderer.sdf_renderer import _on_device_of, _stream

        weight, y = ctx.saved_tensors
        B, Ci, S, U, Co, k = ctx.meta
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        grad_y = grad_y.contiguous()
        lib = _lib.lib()
        with _on_device_of(grad_y):
            g_u = torch.empty((B, Ci, U, U, U), dtype=torch.float32, device=grad_y.device)
            _lib.check(lib.sdfr_conv3d_backward_data(
                grad_y.data_ptr(), None if y is None else y.data_ptr(), B, Ci, U, weight.data_ptr(),
                Co, k, g_u.data_ptr(), _stream()), "sdfr_conv3d_backward_data")
            g_x = g_u
            if S != U:
                g_x = torch.empty((B, Ci, S, S, S), dtype=torch.float32, device=grad_y.device)
                _lib.check(lib.sdfr_upsample3d_backward(g_u.data_ptr(), B * Ci, S, U, g_x.data_ptr(),
                                                        _stream()), "sdfr_upsample3d_backward")
        return g_x, None, None, None, None


def trunk_stage(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], in_size: int,
                relu: bool) -> torch.Tensor:
    """``relu?(conv3d(interpolate(x, in_size, "trilinear", align_corners=False), weight, bias))`` as
    CUDA kernels of ``libsdfrender.so`` (3x3x3 kernels, 4/8/16/32 channels)."""
    return _TrunkStage.apply(x.contiguous(), weight, bias, in_size, relu)


class _FrozenLinearReLU(torch.autograd.Function):
    """``relu(x @ W^T + b)`` of a frozen fully-connected layer whose backward splits the long
    contraction.  The decoder's last fc layer maps 50 -> 8192 features (mug.yaml:2-4), so its data
    gradient is a (B x 8192) @ (8192 x 50) product: one output tile, 8192 deep -- cuBLAS runs it as a
    single-CTA-per-tile kernel in 61 us for B = 64 (profiles/r01q_loop_ops_fused_iteration.txt).
    Cutting the 8192 into ``_SPLIT`` independent slices (one strided-batched GEMM) and adding the
    partial products takes ~8 us.  Still cuBLAS: the fc stack is not a kernel target."""

    _SPLIT = 32

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = torch.relu(torch.addmm(bias, x, weight.t()))
        ctx.save_for_backward(y, weight)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        y, weight = ctx.saved_tensors
        g = grad_y * (y > 0)
        B, K = g.shape
        S = _FrozenLinearReLU._SPLIT
        parts = torch.bmm(g.view(B, S, K // S).transpose(0, 1), weight.view(S, K // S, weight.shape[1]))
        return parts.sum(0), None, None


def _fc_relu(layer: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    if (x.is_cuda and x.requires_grad and layer.bias is not None and not layer.weight.requires_grad
            and layer.out_features >= 2048 and layer.out_features % _FrozenLinearReLU._SPLIT == 0
            and layer.weight.is_contiguous()):
        return _FrozenLinearReLU.apply(x, layer.weight, layer.bias)
    return torch.relu(layer(x))


_CUDA_TRUNK_CHANNELS = (4, 8, 16, 32)


def _cuda_trunk_supported(conv, info) -> bool:
    for layer, (size, _) in zip(conv[:-1], info[:-1]):
        if (tuple(layer.kernel_size) != (3, 3, 3) or tuple(layer.stride) != (1, 1, 1)
                or tuple(layer.padding) != (0, 0, 0) or tuple(layer.dilation) != (1, 1, 1)
                or layer.groups != 1 or layer.in_channels not in _CUDA_TRUNK_CHANNELS
                or layer.out_channels not in _CUDA_TRUNK_CHANNELS or not (3 <= size <= 128)):
            return False
    return True


def _decoder_parts(decoder: nn.Module):
    """(fc layers, conv layers, conv infos as (in_size, relu, kernel, out_channels), volume size)
    of either this package's ``SDFDecoder`` or the reference's (sdfest/vae/sdf_vae.py:170-259:
    ``_fc_layers``, ``_conv_layers``, ``_conv_info`` dicts, ``_volume_size``)."""
    if hasattr(decoder, "_conv_layers"):  # reference class
        info = [(d["in_size"], bool(d["relu"])) for d in decoder._conv_info]
        return list(decoder._fc_layers), list(decoder._conv_layers), info, int(decoder._volume_size)
    info = [(size, bool(relu)) for size, _, _, _, relu in decoder.conv_info]
    return list(decoder.fc), list(decoder.conv), info, int(decoder.volume_size)


class FusedTailDecoder(nn.Module):
    """A frozen SDF decoder whose last stage runs as the fused CUDA tail.

    Wraps this package's ``SDFDecoder`` or the reference's ``sdfest.vae.sdf_vae.SDFDecoder``
    (same weights, no copy).  Everything up to the last convolution stage stays on
    PyTorch/cuDNN; the last stage -- interpolate to the grid resolution + 1x1x1 convolution to one
    channel, 4.8 of the 6.5 ms the whole decoder forward+backward takes for 64 hypotheses on a
    B200 (profiles/r01g_decoder_ops.txt) -- is ``decoder_tail``.  Returns (B,1,R,R,R) like the
    reference.  Raises at construction when the architecture does not end that way.

    ``base`` (R,R,R): optional fixed grid added to the decoded one (a randomly initialised decoder
    emits a near-constant field without a surface; the benchmark decodes residuals around
    an analytic shape).  ``channels_last=True`` stores the trunk's convolution weights in
    ``torch.channels_last_3d``, which makes cuDNN pick its NDHWC tensor-core kernels (2.6 instead of
    4.7 ms for the trunk forward+backward of 64 hypotheses on a B200).

    ``trunk``: ``"cuda"`` runs the convolution stages before the tail through this package's own
    kernels too (``trunk_stage``: every decoder the reference ships is 3x3x3 convolutions between
    4..32 channels with trilinear resizes in between); ``"torch"`` leaves them on PyTorch/cuDNN;
    ``"auto"`` (default) picks ``"cuda"`` when the architecture qualifies.  The fully-connected
    layers stay on PyTorch/cuBLAS either way.  ``trunk_impl`` tells which one runs.
    """

    def __init__(self, decoder: nn.Module, base: Optional[torch.Tensor] = None,
                 trunk: str = "auto", channels_last: bool = True):
        super().__init__()
        self.decoder = decoder
        self.register_buffer("base", None if base is None else base.detach().clone().contiguous())
        fc, conv, info, volume = _decoder_parts(decoder)
        last = conv[-1]
        if (tuple(last.kernel_size) != (1, 1, 1) or last.out_channels != 1 or info[-1][1]
                or info[-1][0] != volume or tuple(last.stride) != (1, 1, 1)):
            raise ValueError("FusedTailDecoder needs a last stage of interpolate-to-volume-size + "
                             "Conv3d(C, 1, kernel_size=1) without ReLU")
        for p in decoder.parameters():
            p.requires_grad_(False)
        self._fc, self._conv, self._info, self.volume_size = fc, conv, info, volume
        if base is not None and tuple(base.shape) != (volume,) * 3:
            raise ValueError(f"base must have shape {(volume,) * 3}, got {tuple(base.shape)}")
        if trunk not in ("auto", "cuda", "torch"):
            raise ValueError("trunk must be 'auto', 'cuda' or 'torch'")
        supported = _cuda_trunk_supported(conv, info)
        if trunk == "cuda" and not supported:
            raise ValueError("trunk='cuda' needs 3x3x3, stride-1, unpadded convolutions between "
                             "4/8/16/32 channels")
        self.trunk_impl = "cuda" if (trunk != "torch" and supported) else "torch"
        if channels_last and self.trunk_impl == "torch":
            for layer in conv[:-1]:
                layer.to(memory_format=torch.channels_last_3d)

    def trunk(self, z: torch.Tensor) -> torch.Tensor:
        """Everything before the last stage's interpolation: (B,L) -> (B,C,S,S,S)."""
        out = z
        for layer in self._fc:
            out = _fc_relu(layer, out)
        c0, s0 = self._conv[0].in_channels, self._info[0][0]
        out = out.view(-1, c0, s0, s0, s0)
        if self.trunk_impl == "cuda" and out.is_cuda:
            for (size, relu), layer in zip(self._info[:-1], self._conv[:-1]):
                out = trunk_stage(out, layer.weight, layer.bias, size, relu)
            return out
        for (size, relu), layer in zip(self._info[:-1], self._conv[:-1]):
            if out.shape[2] != size:
                out = nn.functional.interpolate(out, size=(size,) * 3, mode="trilinear",
                                                align_corners=False)
            out = layer(out)
            if relu:
                out = torch.relu(out)
        return out

    def tail_parameters(self):
        last = self._conv[-1]
        return last.weight.reshape(-1), last.bias

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        w, b = self.tail_parameters()
        return decoder_tail(self.trunk(z), w, b, self.volume_size, self.base)[:, None]
