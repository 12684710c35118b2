"""Batched generation of rendered SDF views.

Reference: ``SDFVAEViewDataset`` (sdfest/initialization/datasets/generated_dataset.py:21-342) feeds
the training of the initialisation network with one sample per ``__next__``: sample a latent,
decode it, sample a pose, call ``render_depth_gpu`` once, post-process on the default stream --
about 10 tiny launches and a dozen host round trips per 640x480 sample.  Here ``batch`` samples are
produced per call: one decode of (batch, L) latents, ONE batched render launch
(``render_depth_batched`` -> ``sdfr_forward``), batched noise models, and a single host
synchronisation when point clouds are requested (their lengths differ per sample).

Same configuration keys and defaults as the reference (generated_dataset.py:97-116); same sampling
distributions, including the reference's x-range ``uniform(-width/2, height/2)``
(generated_dataset.py:264, sic).  Randomness comes from a ``torch.Generator`` instead of Python's
``random`` module, so batches are reproducible per seed but not sample-for-sample identical to the
reference's stream.  The discretised orientation representation (SO3Grid) is out of scope.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterator, List, Optional

import torch

from ..differentiable_renderer import Camera, render_depth_batched
from .losses import depth_to_pointcloud

DEFAULT_CONFIG = {  # generated_dataset.py:97-116
    "width": 640, "height": 480, "fov_deg": 90, "render_threshold": 0.004, "normalize_pose": None,
    "orientation_repr": "quaternion", "mask_noise": False, "mask_noise_min": 0.1,
    "mask_noise_max": 2.0, "norm_noise": False, "norm_noise_min": -0.2, "norm_noise_max": 0.2,
    "scale_to_unit_ball": False, "gaussian_noise_probability": 0.0,
    "gaussian_noise_kernel_size": 5, "gaussian_noise_kernel_std": 1, "pointcloud": False,
}
REQUIRED = ("z_min", "z_max", "extent_mean", "extent_std")


def gaussian_kernel(size: int, std: float, device) -> torch.Tensor:
    """(1,1,size,size) normalised Gaussian (generated_dataset.py:_create_gaussian_kernel)."""
    ax = torch.arange(size, dtype=torch.float32, device=device) - (size - 1) / 2.0
    g = torch.exp(-0.5 * (ax / std) ** 2)
    k = g[:, None] * g[None, :]
    return (k / k.sum())[None, None]


def sample_poses(n: int, camera: Camera, cfg: dict, generator: torch.Generator, device):
    """Positions with the box centre inside the frustum, Shoemake-uniform quaternions, half-extent
    scales (generated_dataset.py:262-272, 196-208, 166-168)."""
    u = torch.rand(n, 6, generator=generator, device=device)
    z = cfg["z_min"] + (cfg["z_max"] - cfg["z_min"]) * u[:, 0]
    x_pix = -camera.width / 2 + (camera.height / 2 + camera.width / 2) * u[:, 1]  # sic, :264
    y_pix = -camera.height / 2 + camera.height * u[:, 2]
    position = torch.stack([x_pix / camera.fx * z, y_pix / camera.fy * z, -z], 1)
    u1, u2, u3 = u[:, 3], u[:, 4], u[:, 5]
    a, b = torch.sqrt(1 - u1), torch.sqrt(u1)
    quat = torch.stack([a * torch.sin(2 * math.pi * u2), a * torch.cos(2 * math.pi * u2),
                        b * torch.sin(2 * math.pi * u3), b * torch.cos(2 * math.pi * u3)], 1)
    scale = (cfg["extent_mean"] + cfg["extent_std"]
             * torch.randn(n, generator=generator, device=device)) / 2.0
    return position.contiguous(), quat.contiguous(), scale.contiguous()


class BatchedSDFViewGenerator:
    """Iterable of batches ``dict(depth (B,H,W), latent_shape (B,L), position (B,3), quaternion
    (B,4), orientation (B,4), scale (B,)[, pointset: list of (N_b,3)])``.

    ``decode``: latents (B,L) -> grids (B,1,R,R,R) or (B,R,R,R) (``SDFVAE.decode``, a
    ``FusedTailDecoder`` ...); ``sample_latent``: n -> (n,L) (default N(0,1), as ``SDFVAE.sample``).
    Samples without a single hit pixel are re-drawn like the reference does (:219-231), batched.
    """

    def __init__(self, config: dict, decode: Callable, latent_size: int, batch: int, device,
                 seed: int = 0, sample_latent: Optional[Callable] = None,
                 render: Callable = render_depth_batched):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config)
        for k in REQUIRED:
            if k not in cfg:
                raise KeyError(f"view dataset config needs {k!r}")
        if cfg["orientation_repr"] != "quaternion":
            raise NotImplementedError("only orientation_repr='quaternion' is supported")
        self.cfg, self.decode, self.latent_size, self.batch = cfg, decode, latent_size, int(batch)
        self.device = torch.device(device)
        f = cfg["width"] / math.tan(cfg["fov_deg"] * math.pi / 180.0 / 2.0) / 2
        self.camera = Camera(cfg["width"], cfg["height"], f, f, cfg["width"] / 2, cfg["height"] / 2,
                             pixel_center=0.5)
        self.generator = torch.Generator(self.device).manual_seed(seed)
        self._sample_latent = sample_latent
        self._render = render
        self._kernel = gaussian_kernel(cfg["gaussian_noise_kernel_size"],
                                       cfg["gaussian_noise_kernel_std"], self.device)

    def __iter__(self) -> Iterator[Dict]:
        while True:
            yield self.generate()

    def _draw(self, n: int):
        if self._sample_latent is not None:
            latent = self._sample_latent(n)
        else:
            latent = torch.randn(n, self.latent_size, generator=self.generator, device=self.device)
        with torch.no_grad():
            sdf = self.decode(latent)
        sdf = sdf[:, 0] if sdf.dim() == 5 else sdf
        position, quat, scale = sample_poses(n, self.camera, self.cfg, self.generator, self.device)
        with torch.no_grad():
            depth = self._render(sdf.contiguous(), position, quat, (1.0 / scale).contiguous(),
                                 self.cfg["render_threshold"], self.camera)
        return latent, position, quat, scale, depth

    def generate(self) -> Dict:
        B = self.batch
        latent, position, quat, scale, depth = self._draw(B)
        for _ in range(16):  # re-draw empty views (one sync per round; rare by construction)
            empty = depth.flatten(1).amax(1) == 0
            n_bad = int(empty.sum())
            if n_bad == 0:
                break
            l2, p2, q2, s2, d2 = self._draw(n_bad)
            idx = torch.nonzero(empty)[:, 0]
            latent[idx], position[idx], quat[idx], scale[idx], depth[idx] = l2, p2, q2, s2, d2
        depth = self.postprocess(depth)
        sample = {"depth": depth, "latent_shape": latent, "position": position, "quaternion": quat,
                  "orientation": quat, "scale": scale}
        if self.cfg["pointcloud"]:
            sample["pointset"] = self.pointsets(sample)
        return sample

    def postprocess(self, depth: torch.Tensor) -> torch.Tensor:
        """Mask noise and Gaussian depth noise (generated_dataset.py:283-307), batched."""
        cfg, gen = self.cfg, self.generator
        B = depth.shape[0]
        exact = depth != 0
        final = exact
        if cfg["mask_noise"]:
            final = self.perturb_masks(exact)
            fill = cfg["mask_noise_min"] + (cfg["mask_noise_max"] - cfg["mask_noise_min"]) * torch.rand(
                B, 1, 1, generator=gen, device=depth.device)
            depth = torch.where(exact, depth, fill.expand_as(depth))
        if cfg["gaussian_noise_probability"] > 0.0:
            apply = torch.rand(B, generator=gen, device=depth.device) < cfg["gaussian_noise_probability"]
            nan_depth = torch.where(depth == 0, torch.full_like(depth, float("nan")), depth)
            filt = torch.nn.functional.conv2d(nan_depth[:, None], self._kernel, padding="same")[:, 0]
            ok = ~(filt.isnan() | filt.isinf())
            blurred = torch.where(ok, filt, nan_depth)
            blurred = torch.nan_to_num(blurred, nan=0.0)
            depth = torch.where(apply[:, None, None], blurred, depth)
        return torch.where(final, depth, torch.zeros_like(depth))

    def perturb_masks(self, masks: torch.Tensor) -> torch.Tensor:
        """Small random affine transform per mask: rotation in [0,1] deg, translation up to 1 % of
        the height vertically, scale in [0.999, 1.001] (torchvision RandomAffine(degrees=(0,1),
        translate=(0.00,0.01), scale=(0.999,1.001)) of generated_dataset.py:242-246), nearest
        sampling, batched through one affine_grid/grid_sample."""
        B, H, W = masks.shape
        gen, dev = self.generator, masks.device
        u = torch.rand(B, 3, generator=gen, device=dev)
        ang = torch.deg2rad(u[:, 0])
        ty = (u[:, 1] * 2 - 1) * 0.01 * 2  # fraction of the height -> normalised [-1,1] units
        s = 0.999 + 0.002 * u[:, 2]
        cos, sin = torch.cos(ang) / s, torch.sin(ang) / s
        theta = torch.zeros(B, 2, 3, device=dev)
        theta[:, 0, 0], theta[:, 0, 1] = cos, -sin * H / W
        theta[:, 1, 0], theta[:, 1, 1], theta[:, 1, 2] = sin * W / H, cos, ty
        grid = torch.nn.functional.affine_grid(theta, (B, 1, H, W), align_corners=False)
        out = torch.nn.functional.grid_sample(masks[:, None].float(), grid, mode="nearest",
                                              padding_mode="zeros", align_corners=False)
        return out[:, 0] > 0.5

    def pointsets(self, sample: Dict) -> List[torch.Tensor]:
        """Per-sample point clouds (OpenGL convention) with the reference's optional
        normalisation (generated_dataset.py:309-333); adjusts position / scale in place."""
        cfg, gen = self.cfg, self.generator
        out = []
        for b in range(sample["depth"].shape[0]):
            pts = depth_to_pointcloud(sample["depth"][b], self.camera)
            if cfg["normalize_pose"]:
                centroid = pts.mean(dim=0)
                pts = pts - centroid
                sample["position"][b] -= centroid
                if cfg["norm_noise"]:
                    noise = cfg["norm_noise_min"] + (cfg["norm_noise_max"] - cfg["norm_noise_min"]) * torch.rand(
                        3, generator=gen, device=pts.device)
                    sample["position"][b] += noise
                    pts = pts + noise
                if cfg["scale_to_unit_ball"]:
                    max_distance = torch.linalg.norm(pts)  # generated_dataset.py:329 (norm of all points, sic)
                    pts = pts / max_distance
                    sample["scale"][b] /= max_distance
            out.append(pts)
        return out
