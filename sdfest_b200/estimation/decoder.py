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
