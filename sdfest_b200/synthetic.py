"""Synthetic workloads: analytic SDF grids (mug / bowl / bottle) and pose-hypothesis sets.

There is no network for ShapeNet grids or trained VAE weights, and a random-init SDFVAE decoder
emits a near-constant field with no iso-surface, so benchmarks, smoke tests
and the demo loop use analytic shapes of the three categories BASELINE.json names.  Grids follow
the renderer's layout: (R,R,R), index order x,y,z, coordinates in [-1,1]^3, y is "up".
"""
from __future__ import annotations

import math

import torch

CATEGORIES = ("mug", "bowl", "bottle")


def _coords(R: int, device, dtype=torch.float32):
    a = torch.linspace(-1.0, 1.0, R, device=device, dtype=dtype)
    return torch.meshgrid(a, a, a, indexing="ij")


def _capped_cylinder(r, y, radius, y0, y1):
    dy = torch.maximum(y0 - y, y - y1)
    dr = r - radius
    return torch.clamp(torch.maximum(dr, dy), max=0.0) + torch.sqrt(
        torch.clamp(dr, min=0.0) ** 2 + torch.clamp(dy, min=0.0) ** 2)


def sdf_mug(R: int = 64, device="cpu", radius: float = 0.42, height: float = 0.62,
            wall: float = 0.07) -> torch.Tensor:
    """Open cylinder with a floor and a torus-segment handle."""
    x, y, z = _coords(R, device)
    r = torch.sqrt(x * x + z * z)
    outer = _capped_cylinder(r, y, radius, -height, height)
    inner = _capped_cylinder(r, y, radius - wall, -height + wall, height + 1.0)
    body = torch.maximum(outer, -inner)
    # handle: torus in the z=0 plane centred on the rim of the body
    hx = x - (radius + 0.12)
    ring = torch.sqrt(hx * hx + y * y) - 0.26
    handle = torch.sqrt(ring * ring + z * z) - 0.055
    handle = torch.maximum(handle, radius - 0.02 - x)  # keep only the part outside the body
    return torch.minimum(body, handle).contiguous()


def sdf_bowl(R: int = 64, device="cpu", radius: float = 0.68, wall: float = 0.06) -> torch.Tensor:
    """Lower hemisphere shell."""
    x, y, z = _coords(R, device)
    shell = torch.abs(torch.sqrt(x * x + y * y + z * z) - radius) - wall
    return torch.maximum(shell, y - 0.12).contiguous()


def sdf_bottle(R: int = 64, device="cpu", radius: float = 0.34, neck: float = 0.13) -> torch.Tensor:
    """Solid body cylinder with a thinner neck."""
    x, y, z = _coords(R, device)
    r = torch.sqrt(x * x + z * z)
    body = _capped_cylinder(r, y, radius, -0.82, 0.3)
    top = _capped_cylinder(r, y, neck, 0.25, 0.82)
    return torch.minimum(body, top).contiguous()


def sdf_sphere(R: int = 64, device="cpu", radius: float = 0.6) -> torch.Tensor:
    x, y, z = _coords(R, device)
    return (torch.sqrt(x * x + y * y + z * z) - radius).contiguous()


def category_grid(category: str, R: int, device, shape_param: float = 0.0) -> torch.Tensor:
    """One grid of a category; ``shape_param`` in [-1,1] varies the shape (a stand-in latent)."""
    k = 1.0 + 0.12 * float(shape_param)
    if category == "mug":
        return sdf_mug(R, device, radius=0.42 * k, height=0.62 / k)
    if category == "bowl":
        return sdf_bowl(R, device, radius=0.68 * min(k, 1.1))
    if category == "bottle":
        return sdf_bottle(R, device, radius=0.34 * k, neck=0.13 * k)
    raise ValueError(f"unknown category {category!r}")


def random_unit_quaternions(n: int, generator: torch.Generator) -> torch.Tensor:
    """Shoemake-uniform unit quaternions (x,y,z,w) -- recipe of estimation/simple_setup.py:856-868."""
    u = torch.rand(n, 3, generator=generator)
    a, b = torch.sqrt(1 - u[:, 0]), torch.sqrt(u[:, 0])
    t1, t2 = 2 * math.pi * u[:, 1], 2 * math.pi * u[:, 2]
    return torch.stack([a * torch.sin(t1), a * torch.cos(t1), b * torch.sin(t2), b * torch.cos(t2)], 1)


def quaternion_multiply(q1: torch.Tensor, q2: torch.Tensor) -> torch.Tensor:
    """Hamilton product, scalar-last (initialization/quaternion_utils.py:12-38)."""
    x1, y1, z1, w1 = q1.unbind(-1)
    x2, y2, z2, w2 = q2.unbind(-1)
    return torch.stack([
        w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], -1)


def make_hypotheses(n: int, seed: int = 0, device="cpu", base_position=(0.02, -0.01, -0.4),
                    base_scale: float = 0.15, pos_sigma: float = 0.03, rot_deg: float = 10.0,
                    scale_rel: float = 0.10):
    """The hypothesis set of BASELINE config 2: a base pose perturbed by
    N(0, pos_sigma) position noise, rotations of up to ``rot_deg`` degrees about random axes
    and +-``scale_rel`` relative scale.  Hypothesis 0 is the unperturbed pose.

    Returns dict(position (n,3), orientation (n,4), inv_scale (n,), shape_param (n,)) float32.
    """
    g = torch.Generator().manual_seed(seed)
    base_q = random_unit_quaternions(1, torch.Generator().manual_seed(1))
    pos = torch.tensor(base_position).repeat(n, 1) + pos_sigma * torch.randn(n, 3, generator=g)
    axis = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    ang = (torch.rand(n, generator=g) * 2 - 1) * math.radians(rot_deg)
    dq = torch.cat([axis * torch.sin(ang / 2)[:, None], torch.cos(ang / 2)[:, None]], 1)
    quat = torch.nn.functional.normalize(quaternion_multiply(dq, base_q.expand(n, 4)), dim=1)
    scale = base_scale * (1 + scale_rel * (torch.rand(n, generator=g) * 2 - 1))
    shape = torch.rand(n, generator=g) * 2 - 1
    pos[0] = torch.tensor(base_position)
    quat[0] = base_q[0]
    scale[0] = base_scale
    shape[0] = 0.0
    return {
        "position": pos.float().contiguous().to(device),
        "orientation": quat.float().contiguous().to(device),
        "inv_scale": (1.0 / scale).float().contiguous().to(device),
        "shape_param": shape.float().contiguous().to(device),
    }


def hypothesis_grids(shape_params, R: int, device, categories=("mug",)) -> torch.Tensor:
    """(n,R,R,R) grids: hypothesis i has category ``categories[i % len]`` and its own shape."""
    params = [float(v) for v in shape_params.tolist()]
    out = torch.empty(len(params), R, R, R, dtype=torch.float32, device=device)
    for i, p in enumerate(params):
        out[i] = category_grid(categories[i % len(categories)], R, "cpu", p)
    return out


def residual_decoder(R: int, device, base: torch.Tensor, seed: int = 0, gain: float = 1.0,
                     trunk: str = "auto"):
    """A decoder of the reference's architecture that always has a surface: a randomly initialised
    ``SDFDecoder`` (frozen) decoding RESIDUALS around the analytic grid ``base`` through the fused
    CUDA tail (``estimation.FusedTailDecoder``).  The last convolution is rescaled by ``gain`` and its
    bias re-centred so that the residual of latent 0 has zero mean -- a random-init decoder emits a
    near-constant field (0.37 everywhere for seed 0) that would otherwise push the
    surface out of the grid."""
    from .estimation.decoder import FusedTailDecoder, SDFDecoder

    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    dec = SDFDecoder(R).to(device).eval()
    torch.random.set_rng_state(gen_state)
    with torch.no_grad():
        last = dec.conv[-1]
        last.weight.mul_(gain)
        last.bias.mul_(gain)
        mean0 = dec(torch.zeros(1, dec.fc[0].in_features, device=device)).mean()
        last.bias.sub_(mean0)
    return FusedTailDecoder(dec, base=base.to(device), trunk=trunk).to(device).eval()
