"""Shared scene builders for the test-suite (analytic SDFs, cameras, golden loader)."""
from __future__ import annotations

import glob
import math
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POSE_KEYS = ("x", "y", "z", "qx", "qy", "qz", "qw", "s_inv")


def golden_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*_r*.npz")))


def load_golden(name):
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    if "sdf" not in z:
        z["sdf"] = np.load(os.path.join(GOLDEN_DIR, str(z["sdf_file"])))["sdf"]
    W, H = int(z["width"]), int(z["height"])
    f = W / math.tan(float(z["fov_deg"]) * math.pi / 180.0 / 2.0) / 2  # sdf_renderer.py:419
    z["cam"] = dict(cx=W / 2, cy=H / 2, fx=f, fy=f)
    z["W"], z["H"] = W, H
    return z


def mug_sdf():
    return np.load(os.path.join(GOLDEN_DIR, "mug_z0_sdf.npz"))["sdf"]


def grid_coords(R):
    a = np.linspace(-1.0, 1.0, R)
    return np.meshgrid(a, a, a, indexing="ij")


def sdf_sphere(R, r=0.6):
    x, y, z = grid_coords(R)
    return (np.sqrt(x * x + y * y + z * z) - r).astype(np.float32)


def sdf_torus(R, major=0.55, minor=0.2):
    x, y, z = grid_coords(R)
    q = np.sqrt(x * x + z * z) - major
    return (np.sqrt(q * q + y * y) - minor).astype(np.float32)


def sdf_box(R, half=(0.5, 0.35, 0.6)):
    x, y, z = grid_coords(R)
    qx, qy, qz = np.abs(x) - half[0], np.abs(y) - half[1], np.abs(z) - half[2]
    outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
    return (outside + np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0)).astype(np.float32)


def sdf_bottle(R):
    """Capped cylinder body + thinner neck (the 'bottle' of BASELINE config 4)."""
    x, y, z = grid_coords(R)
    r = np.sqrt(x * x + z * z)

    def capped(rad, y0, y1):
        dy = np.maximum(y0 - y, y - y1)
        dr = r - rad
        return np.minimum(np.maximum(dr, dy), 0) + np.sqrt(np.maximum(dr, 0) ** 2 + np.maximum(dy, 0) ** 2)

    return np.minimum(capped(0.35, -0.8, 0.3), capped(0.14, 0.25, 0.8)).astype(np.float32)


def sdf_bowl(R):
    """Hemispherical shell (the 'bowl' of BASELINE config 4)."""
    x, y, z = grid_coords(R)
    shell = np.abs(np.sqrt(x * x + y * y + z * z) - 0.65) - 0.06
    return np.maximum(shell, y - 0.1).astype(np.float32)


def shoemake(seed):
    """Uniform random unit quaternion (x,y,z,w) -- recipe of estimation/simple_setup.py:856-868."""
    u1, u2, u3 = np.random.default_rng(seed).random(3)
    return np.array([
        np.sqrt(1 - u1) * np.sin(2 * np.pi * u2),
        np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
        np.sqrt(u1) * np.sin(2 * np.pi * u3),
        np.sqrt(u1) * np.cos(2 * np.pi * u3),
    ], dtype=np.float32)


def default_camera(W=640, H=480):
    """estimation/configs/default.yaml:1-8 scaled to (W,H): f = W/2, principal point centred."""
    return dict(cx=W / 2, cy=H / 2, fx=W / 2, fy=W / 2)


def depth_parity(depth, ref, threshold, rtol=1e-5, frac=0.999):
    """The depth gate of BASELINE.md: >= `frac` of the pixels hit by either renderer agree to
    `rtol` relative (with identical hit/miss state); the remainder -- sphere-trace termination
    flipped by one step under rounding -- must stay within `threshold` relative, or be a hit/miss
    flip.  Returns a dict of diagnostics and asserts the gate."""
    depth = np.asarray(depth, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    either = (depth != 0) | (ref != 0)
    both = (depth != 0) & (ref != 0)
    n = int(either.sum())
    rel = np.zeros_like(ref)
    rel[both] = np.abs(depth[both] - ref[both]) / np.abs(ref[both])
    good = both & (rel <= rtol)
    n_good = int(good.sum())
    mask_flips = int((either & ~both).sum())
    loose = both & (rel > rtol)
    info = dict(n=n, good=n_good, mask_flips=mask_flips, step_flips=int(loose.sum()),
                max_rel_good=float(rel[good].max()) if n_good else 0.0,
                bit_equal=float((depth[either] == ref[either]).mean()) if n else 1.0)
    assert n == 0 or n_good >= frac * n, info
    if loose.any():
        assert rel[loose].max() <= 4 * threshold + 1e-3, info
    return info


def grad_close(a, b, rtol=1e-3, what=""):
    """Gradient gate: |a-b| <= rtol * max|b| elementwise (fp32 atomics reorder the sums)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{what}: max err {err:.3e} relative to max|ref|={scale:.3e}"
    return err


def sdf_grad_parity(a, b, rtol=1e-3, pixel_bound=None, max_flip_pixels=4, what=""):
    """SDF-gradient gate.  |a-b| <= rtol * max|b| for all voxels, EXCEPT the voxels of at most
    `max_flip_pixels` hit pixels whose cell lookup flipped: the backward re-derives the hit point from the
    stored depth and floors it to a cell (sdf_renderer_cuda.cu:336-354); a coordinate within an ulp of a
    cell face lands in either cell depending on how the compiler contracts `o + t*d` into an fma, and the
    reference's weight list (cu:373-388) is not continuous across cell faces, so such a pixel moves its
    whole contribution (8 + 8 voxels).  Each outlier is bounded by one pixel's contribution `pixel_bound`
    (max |grad_depth| * scale).  Returns (max relative error of the inliers, number of outlier voxels)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b)
    bad = err > rtol * scale
    n_bad = int(bad.sum())
    assert n_bad <= 16 * max_flip_pixels, f"{what}: {n_bad} voxels off by more than {rtol} of max|ref|={scale:.3e}"
    if n_bad:
        bound = pixel_bound if pixel_bound is not None else 0.2 * scale
        assert err[bad].max() <= 1.001 * bound, f"{what}: outlier {err[bad].max():.3e} exceeds one pixel's contribution {bound:.3e}"
    return float(err[~bad].max() / scale), n_bad


def pose_grad_parity(got, bw, g, rtol=1e-3, what="", position=None, inv_scale=None):
    """Pose-gradient gate against an oracle backward `bw` run with want_deriv=True and upstream image
    `g`: per parameter group (position, orientation, inv_scale)

        |got - ref| <= rtol * max|ref| + 16 eps * kappa * S_group,        eps = 2^-24.

    The first term is BASELINE.json's 1e-3.  The second is the forward error bound of ANY fp32
    evaluation of these sums, which takes over when a gradient (nearly) cancels and rtol of its tiny
    maximum asks for digits fp32 does not have -- the reference's own kernel does not reproduce itself to
    that either (its fp32 oracle and its float64 oracle differ by 4e-3 of the maximum on the C3 scene):
      S_group  = sum over the hit pixels of |g * d depth / d theta|, largest component of the group; for
                 the orientation additionally 2 sqrt(3) scale * S_position -- the size of the two halves
                 of d c / d q (cu:402-437: C_k (x - p) and -2 q_k o) BEFORE they cancel, which is what
                 bounds the rounding of every formulation, per pixel (reference) or per moment (ours),
                 when the shape is rotationally symmetric and the halves cancel exactly;
      kappa    = max(1, |p| inv_scale): the hit point in the object frame is the difference of two
                 vectors of length |p| (cu:344-345), so its rounding is |p| / scale relative.
    `got` = (position (3,), orientation (4,), inv_scale scalar).  Returns the largest error relative to
    rtol * max|ref| (<= 1 means inside the first term alone)."""
    ref = (np.asarray(bw["g_position"], np.float64), np.asarray(bw["g_orientation"], np.float64),
           np.asarray([bw["g_inv_scale"]], np.float64))
    absum = np.abs(np.asarray(g, np.float64)[None] * np.asarray(bw["deriv"], np.float64)).sum((1, 2))
    kappa, scale_obj = 1.0, 0.0
    if position is not None and inv_scale is not None:
        scale_obj = 1.0 / float(inv_scale)
        kappa = max(1.0, float(np.linalg.norm(np.asarray(position, np.float64))) * float(inv_scale))
    s_group = (absum[0:3].max(), max(absum[3:7].max(), 2 * 3 ** 0.5 * scale_obj * absum[0:3].max()), absum[7])
    worst = 0.0
    for a, b, sg, nm in zip(got, ref, s_group, ("position", "orientation", "inv_scale")):
        a = np.asarray(a, np.float64).reshape(-1)
        scale = max(np.abs(b).max(), 1e-30)
        err = np.abs(a - b).max()
        tol = rtol * scale + 16 * 2.0 ** -24 * kappa * sg
        assert err <= tol, f"{what} {nm}: err {err:.3e} > {tol:.3e} (max|ref| {scale:.3e}, S {sg:.3e}, kappa {kappa:.2f})"
        worst = max(worst, err / (rtol * scale))
    return worst


def cell_face_mask(depth, position, orientation, inv_scale, R, cx, cy, fx, fy, delta=2e-3):
    """Hit pixels whose hit point (re-derived from the depth as cu:336-354 does, here in float64) lies
    within `delta` cells of a cell face.  There the reference's cell lookup `floor` can go either way under
    fp32 rounding, and both the corner weights (cu:373-388) and the trilinear gradient that carries the pose
    derivatives (cu:444-456) are DISCONTINUOUS across the face -- one such pixel moves the gradients by its
    whole contribution.  Tests with an explicit upstream image zero it on these pixels (about 1 % of the
    hits), so that renderer and oracle are compared where the function they differentiate is smooth."""
    z = np.asarray(depth, np.float64)
    H, W = z.shape
    col, row = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    d = np.stack([(col + 0.5 - cx) / fx, -(row + 0.5 - cy) / fy, -np.ones_like(col)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    t = -z / d[..., 2]
    x, y, zq, w = (float(v) for v in np.asarray(orientation, np.float64).reshape(-1)[:4])
    Rm = np.array([[1 - 2 * (y * y + zq * zq), 2 * (x * y - w * zq), 2 * (x * zq + w * y)],
                   [2 * (x * y + w * zq), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - w * x)],
                   [2 * (x * zq - w * y), 2 * (y * zq + w * x), 1 - 2 * (x * x + y * y)]])
    o = (t[..., None] * d - np.asarray(position, np.float64).reshape(-1)[:3]) @ Rm  # R^T (x - p)
    v = (o * float(np.asarray(inv_scale).reshape(-1)[0]) + 1.0) * (R - 1) / 2.0
    frac = v - np.floor(v)
    near = (np.minimum(frac, 1.0 - frac) < delta).any(-1)
    return near & (z != 0)
