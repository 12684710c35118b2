"""torch restatement of the reference's point-cloud loss -- TEST INFRASTRUCTURE ONLY.

Reference: sdfest/estimation/losses.py:32-135 (``pc_loss``: trilinear SDF lookup at observed points,
0 outside the volume), as used by estimation/simple_setup.py:134-144 (``loss_pc = mean |pc_loss|``).
Batched over hypotheses; pinned to the reference's own outputs and autograd gradients by
tests/golden/pcloss_torus16.npz (tests/test_estimation_host.py).  The product evaluates this loss with
the kernels of libsdfrender.so only; the CPU tests of the host logic patch this in as the stand-in.
"""
from __future__ import annotations

import torch


def rotation_matrices(q: torch.Tensor) -> torch.Tensor:
    """(B,4) unit quaternions (x,y,z,w) -> (B,3,3) matrices of the INVERSE rotation, i.e. camera
    -> object, laid out as in losses.py:61-75."""
    x, y, z, w = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - w * y),
        2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x),
        2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)


def pc_loss(points: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor,
            scale: torch.Tensor, sdf: torch.Tensor) -> torch.Tensor:
    """Trilinearly interpolated SDF value at observed points, batched over hypotheses.

    points (M,3); position (B,3); orientation (B,4); scale (B,); sdf (B,R,R,R) or (1,R,R,R).
    Returns (B,M) distances in world units, 0 for points outside the SDF volume
    (losses.py:84-135).  Differentiable w.r.t. position, orientation, scale and sdf.
    """
    B = position.shape[0]
    res = sdf.shape[-1]
    q = orientation / torch.linalg.norm(orientation, dim=1, keepdim=True)  # losses.py:56
    obj = torch.einsum("bij,bmj->bmi", rotation_matrices(q), points[None] - position[:, None])
    obj = obj / scale[:, None, None]
    grid_size = 2.0 / (res - 1)
    c = torch.floor((obj + 1.0) * (res - 1) * 0.5)
    outside = (c.min(dim=2)[0] < 0) | (c.max(dim=2)[0] > res - 2)
    c = torch.clamp(c, 0, res - 2)
    off = (obj - (c * grid_size - 1.0)) / grid_size
    ci = c.long()
    base = (ci[..., 0] * res + ci[..., 1]) * res + ci[..., 2]  # (B,M)
    flat = sdf.reshape(sdf.shape[0], -1).expand(B, -1)

    def corner(dx, dy, dz):
        return torch.gather(flat, 1, base + (dx * res + dy) * res + dz)

    ox, oy, oz = off.unbind(-1)
    val = ((corner(0, 0, 0) * (1 - ox) + corner(1, 0, 0) * ox) * (1 - oy)
           + (corner(0, 1, 0) * (1 - ox) + corner(1, 1, 0) * ox) * oy) * (1 - oz) \
        + ((corner(0, 0, 1) * (1 - ox) + corner(1, 0, 1) * ox) * (1 - oy)
           + (corner(0, 1, 1) * (1 - ox) + corner(1, 1, 1) * ox) * oy) * oz
    val = torch.where(outside, torch.zeros_like(val), val)
    return val * scale[:, None]


def point_loss(points: torch.Tensor, position: torch.Tensor, orientation: torch.Tensor,
               scale: torch.Tensor, sdf: torch.Tensor) -> torch.Tensor:
    """Per-hypothesis ``mean_m |pc_loss[b, m]|``; points (M,3) or (B,M,3); sdf (B|1,R,R,R)."""
    if points.dim() == 3:
        return torch.stack([pc_loss(points[b], position[b:b + 1], orientation[b:b + 1],
                                    scale[b:b + 1], sdf[b:b + 1] if sdf.shape[0] > 1 else sdf)[0]
                            for b in range(position.shape[0])]).abs().mean(dim=1)
    return pc_loss(points, position, orientation, scale, sdf).abs().mean(dim=1)
