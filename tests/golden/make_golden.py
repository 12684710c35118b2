"""Generate golden vectors from the REFERENCE's own CPU renderer.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It imports ``sdfest/differentiable_renderer/simple_renderer.py`` by file path through
three shims (stub ``matplotlib``, ``np.float = float`` -- removed in numpy>=1.24 but used
at simple_renderer.py:271,275 -- and no package import, which would pull open3d and a
JIT CUDA build), renders a handful of small seeded scenes with
``render_depth(..., value="d" | "c")`` and stores inputs + outputs as ``*.npz`` next to
this script.  ``mug_z0`` uses the trained VAE shipped with the reference's tests
(``tests/initilization/vae_model/mug.pt``) decoded at z = 0.

The reference cannot travel to the GPU box, the fixtures do: tests never import the
reference, they only read these files.

Stored per scene (float64 unless noted):
  sdf (float32 -- every value is exactly representable, the renderer is fed sdf.astype(f64);
       scenes sharing a big grid store "sdf_file" = name of an .npz holding it instead),
  position, orientation (x,y,z,w), inv_scale, width, height, fov_deg, threshold,
  depth (H,W)  = image returned for value="d",
  steps (H,W)  = image returned for value="c" (trilinear samples per hit ray; 0 where no hit),
  deriv (8,H,W)= derivatives[k] for k in x,y,z,qx,qy,qz,qw,s_inv,
  g (H,W)      = seeded upstream gradient,
  g_sdf_exact (R,R,R) = sum_pix g * derivatives["sdf"][voxel]   (simple_renderer.py:399-408),
  g_pose (8,)  = sum_pix g * derivatives[k]                      (sdf_renderer.py:250-257).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
POSE_KEYS = ("x", "y", "z", "qx", "qy", "qz", "qw", "s_inv")


def load_reference_renderer():
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    if not hasattr(np, "float"):
        np.float = float  # simple_renderer.py:271,275
    spec = importlib.util.spec_from_file_location(
        "ref_simple_renderer",
        os.path.join(REF, "sdfest/differentiable_renderer/simple_renderer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def grid_coords(R):
    a = np.linspace(-1.0, 1.0, R)
    return np.meshgrid(a, a, a, indexing="ij")


def sdf_sphere(R, r=0.6):
    x, y, z = grid_coords(R)
    return np.sqrt(x * x + y * y + z * z) - r


def sdf_torus(R, major=0.55, minor=0.2):
    x, y, z = grid_coords(R)
    q = np.sqrt(x * x + z * z) - major
    return np.sqrt(q * q + y * y) - minor


def sdf_box(R, half=(0.5, 0.35, 0.6)):
    x, y, z = grid_coords(R)
    qx, qy, qz = np.abs(x) - half[0], np.abs(y) - half[1], np.abs(z) - half[2]
    outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
    inside = np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0)
    return outside + inside


def sdf_shell(R, r=0.7, thick=0.08):
    """Hollow sphere: the camera can sit inside the object box and inside the cavity."""
    x, y, z = grid_coords(R)
    return np.abs(np.sqrt(x * x + y * y + z * z) - r) - thick


def shoemake(seed):
    """Uniform random unit quaternion (x,y,z,w) (recipe of simple_setup.py:856-868)."""
    u1, u2, u3 = np.random.default_rng(seed).random(3)
    return np.array([
        np.sqrt(1 - u1) * np.sin(2 * np.pi * u2),
        np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
        np.sqrt(u1) * np.sin(2 * np.pi * u3),
        np.sqrt(u1) * np.cos(2 * np.pi * u3),
    ])


def mug_sdf(latent):
    import torch
    import yaml

    sys.path.insert(0, REF)
    from sdfest.vae.sdf_vae import SDFVAE

    cfg = yaml.safe_load(open(os.path.join(REF, "tests/initilization/vae_model/mug.yaml")))
    vae = SDFVAE(sdf_size=64, latent_size=cfg["latent_size"], encoder_dict=cfg["encoder"],
                 decoder_dict=cfg["decoder"], device="cpu", tsdf=cfg["tsdf"])
    state = torch.load(os.path.join(REF, "tests/initilization/vae_model/mug.pt"),
                       map_location="cpu")
    vae.load_state_dict(state)
    vae.eval()
    with torch.no_grad():
        return vae.decode(torch.as_tensor(latent, dtype=torch.float32)[None])[0, 0].numpy()


def run_scene(sr, name, sdf, position, orientation, inv_scale, width, height, fov_deg,
              threshold, seed, sdf_file=None):
    sdf32 = np.ascontiguousarray(sdf, dtype=np.float32)
    sdf64 = sdf32.astype(np.float64)
    position = np.asarray(position, dtype=np.float64)
    orientation = np.asarray(orientation, dtype=np.float64)
    orientation = orientation / np.linalg.norm(orientation)
    inv_scale = float(inv_scale)
    obj = sr.SDFObject(sdf64)
    with contextlib.redirect_stdout(io.StringIO()):
        depth, derivatives = sr.render_depth(obj, width, height, fov_deg, "d", threshold,
                                             position.copy(), orientation.copy(), inv_scale)
        steps, _ = sr.render_depth(obj, width, height, fov_deg, "c", threshold,
                                   position.copy(), orientation.copy(), inv_scale)
    g = np.random.default_rng(seed).standard_normal((height, width))
    deriv = np.stack([np.asarray(derivatives[k]) if k in derivatives
                      else np.zeros((height, width)) for k in POSE_KEYS])
    g_pose = np.array([np.sum(d * g) for d in deriv])
    R = sdf32.shape[0]
    g_sdf = np.zeros((R, R, R))
    if "sdf" in derivatives:
        for idx, img in derivatives["sdf"].items():
            g_sdf[idx] = np.sum(img * g)
    out = os.path.join(HERE, f"{name}.npz")
    # large grids shared by several scenes live in their own file (key "sdf_file")
    grid = {"sdf": sdf32} if sdf_file is None else {"sdf_file": np.str_(sdf_file)}
    np.savez_compressed(
        out, **grid, position=position, orientation=orientation,
        inv_scale=np.float64(inv_scale), width=np.int64(width), height=np.int64(height),
        fov_deg=np.float64(fov_deg), threshold=np.float64(threshold), depth=depth,
        steps=steps.astype(np.int32), deriv=deriv, g=g, g_sdf_exact=g_sdf, g_pose=g_pose)
    hits = int((depth > 0).sum())
    print(f"{name}: {width}x{height} R={R} hits={hits} max_steps={int(steps.max())} "
          f"-> {os.path.getsize(out) / 1024:.0f} KiB")
    assert hits > 0, "golden scene without a single hit is useless"


def pc_loss_fixture():
    """Golden vector for the caller-side point-cloud loss (estimation/losses.py:32-135)."""
    import torch

    sys.path.insert(0, REF)
    from sdfest.estimation.losses import pc_loss

    rng = np.random.default_rng(11)
    sdf = torch.tensor(sdf_torus(16), dtype=torch.float64)
    points = torch.tensor(rng.uniform(-0.33, 0.33, (200, 3)) + np.array([0.05, -0.03, -0.8]))
    pos = torch.tensor([0.05, -0.03, -0.8], dtype=torch.float64, requires_grad=True)
    quat = torch.tensor(shoemake(3) * 1.3, requires_grad=True)  # un-normalised on purpose
    scale = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    sdf.requires_grad_(True)
    val = pc_loss(points, pos, quat, scale, sdf)
    val.abs().mean().backward()
    np.savez_compressed(
        os.path.join(HERE, "pcloss_torus16.npz"), sdf=sdf.detach().numpy(), points=points.numpy(),
        position=pos.detach().numpy(), orientation=quat.detach().numpy(),
        scale=scale.detach().numpy(), value=val.detach().numpy(), g_position=pos.grad.numpy(),
        g_orientation=quat.grad.numpy(), g_scale=scale.grad.numpy(), g_sdf=sdf.grad.numpy())
    print("pcloss_torus16: outside points", int((val == 0).sum().item()), "of", val.numel())


def main():
    pc_loss_fixture()
    sr = load_reference_renderer()
    # 1. sphere, identity pose, coarse grid
    run_scene(sr, "sphere_r16", sdf_sphere(16), [0.0, 0.0, -1.0], [0, 0, 0, 1], 1 / 0.4,
              32, 24, 60.0, 0.005, seed=1)
    # 2. torus, random orientation, tighter threshold, off-centre
    run_scene(sr, "torus_r32", sdf_torus(32), [0.05, -0.03, -0.8], shoemake(1), 1 / 0.3,
              48, 36, 50.0, 0.003, seed=2)
    # 3. camera inside the object's box and inside a hollow shell (t_min = 0 path)
    run_scene(sr, "shell_inside_r24", sdf_shell(24), [0.02, 0.01, -0.05], shoemake(2), 1 / 0.5,
              40, 30, 90.0, 0.01, seed=3)
    # 4. box partially off-screen, rotated by 90 deg about y: rays parallel to slabs at centre
    run_scene(sr, "box_offscreen_r24", sdf_box(24), [0.45, 0.2, -0.9],
              [0.0, np.sin(np.pi / 4), 0.0, np.cos(np.pi / 4)], 1 / 0.35,
              40, 30, 60.0, 0.005, seed=4)
    # 5. trained mug VAE at z = 0 (the reference's own test fixture), default-like view
    mug = mug_sdf(np.zeros(8))
    np.savez_compressed(os.path.join(HERE, "mug_z0_sdf.npz"), sdf=mug.astype(np.float32))
    run_scene(sr, "mug_z0_r64", mug, [0.02, -0.01, -0.4], shoemake(1), 1 / 0.15,
              64, 48, 90.0, 0.005, seed=5, sdf_file="mug_z0_sdf.npz")
    # 6. same grid, different view/threshold
    run_scene(sr, "mug_z0_r64_far", mug, [-0.03, 0.02, -0.6], shoemake(7), 1 / 0.12,
              48, 36, 45.0, 0.003, seed=6, sdf_file="mug_z0_sdf.npz")


if __name__ == "__main__":
    main()
