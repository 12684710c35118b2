"""numpy statement of the empty-space bounds (``sdfr_grid_bounds``) -- TEST INFRASTRUCTURE ONLY.

The reference marches every ray that enters the grid's box (sdf_renderer_cuda.cu:272-293) and stops
at the first sample with ``dist < threshold * t`` (:286), ``dist = trilinear * scale`` (:285).  A sample
can therefore only stop the march where ``trilinear < threshold * t / scale``; ``t`` never exceeds
``|p| + sqrt(3) scale`` inside the box, and a trilinear value is never below the smallest of its cell's
8 corners.  ``cell_bounds`` returns, per axis, the first / last cell whose smallest corner is below that
bound ``tau`` -- a ray that misses the box of those cells cannot hit anything.
"""
from __future__ import annotations

import numpy as np


def hit_tau(position, inv_scale, threshold) -> np.float32:
    """The bound of sdfr_core.cuh::hit_tau, evaluated in float32 in the same order."""
    f = np.float32
    p = np.asarray(position, dtype=np.float32).reshape(-1)[:3]
    inv_scale = f(inv_scale)
    scale = f(1.0 / np.float64(inv_scale))
    far = f(np.sqrt(f(f(f(p[0] * p[0]) + f(p[1] * p[1])) + f(p[2] * p[2])))) + f(f(1.7320509) * scale)
    return f(f(f(f(f(threshold) * f(far)) * inv_scale) * f(1.001)) + f(1e-5))


def cell_min(sdf) -> np.ndarray:
    """(R-1)^3 smallest corner value of every cell."""
    s = np.asarray(sdf, dtype=np.float32)
    m = np.minimum(s[:-1], s[1:])
    m = np.minimum(m[:, :-1], m[:, 1:])
    return np.minimum(m[:, :, :-1], m[:, :, 1:])


def cell_bounds(sdf, tau):
    """(lo (3,), hi (3,)) int cell indices; lo > hi when no cell is below tau."""
    below = cell_min(sdf) < np.float32(tau)
    lo, hi = np.full(3, 0x7FFFFFFF, np.int64), np.full(3, -1, np.int64)
    if below.any():
        for a in range(3):
            other = tuple(i for i in range(3) if i != a)
            idx = np.nonzero(below.any(axis=other))[0]
            lo[a], hi[a] = idx[0], idx[-1]
    return lo, hi


def pack(lo, hi, tau) -> np.ndarray:
    """The 32-byte record of include/sdfrender.h::sdfr_cell_bounds as 8 int32."""
    out = np.zeros(8, np.int32)
    out[0:3] = np.asarray(lo, np.int64).astype(np.int32)
    out[3:6] = np.asarray(hi, np.int64).astype(np.int32)
    out[6] = np.asarray([tau], np.float32).view(np.int32)[0]
    return out
