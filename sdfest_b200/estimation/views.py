"""Several camera views of one object: the rigid maps of the reference's view loop and their adjoints.

Reference: estimation/simple_setup.py:420-431 moves the object pose into every camera frame,
    q_w2c = quaternion_invert(camera_orientation)
    position_c = quaternion_apply(q_w2c, position - camera_position)
    orientation_c = quaternion_multiply(q_w2c, norm_orientation)
(initialization/quaternion_utils.py:12-66; scalar-last quaternions), renders each view and sums the
per-view losses.  Both maps are LINEAR in the object pose: position_c = R_v (p - c_v) with R_v the
rotation matrix of q_w2c, orientation_c = L_v q with L_v the left-multiplication matrix of q_w2c.  So
the renderer's pose gradients, which arrive per view in the camera frame, are pulled back with
R_v^T and L_v^T and summed over the views -- no autograd graph, a handful of batched matrix products
in front of the optimiser step.  Scale and SDF gradients need no pull-back.

This module is the host-side mathematics (checked on CPU against autograd and against golden vectors
of the reference's own quaternion_utils); the loop that drives V renders per hypothesis with it is
listed under "what is next" in DESIGN.md.
"""
from __future__ import annotations

import torch


def quaternion_invert(quaternions: torch.Tensor) -> torch.Tensor:
    """Conjugate (= inverse for unit quaternions), (...,4) scalar-last (quaternion_utils.py:57-66)."""
    return torch.cat([-quaternions[..., :3], quaternions[..., 3:]], -1)


def left_matrix(q: torch.Tensor) -> torch.Tensor:
    """L(q) (...,4,4) with  q (x) r = L(q) r  for scalar-last quaternions (quaternion_utils.py:28-34)."""
    x, y, z, w = q.unbind(-1)
    return torch.stack([
        torch.stack([w, -z, y, x], -1),
        torch.stack([z, w, -x, y], -1),
        torch.stack([-y, x, w, z], -1),
        torch.stack([-x, -y, -z, w], -1)], -2)


def quaternion_multiply(quaternions_1: torch.Tensor, quaternions_2: torch.Tensor) -> torch.Tensor:
    """Hamilton product q1 (x) q2 with broadcasting (quaternion_utils.py:12-34)."""
    return (left_matrix(quaternions_1) @ quaternions_2[..., None])[..., 0]


def rotation_matrix(q: torch.Tensor) -> torch.Tensor:
    """(...,3,3) rotation of a UNIT quaternion: quaternion_apply(q, v) = R(q) v."""
    x, y, z, w = q.unbind(-1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


def quaternion_apply(quaternions: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """q (x) (v,0) (x) conj(q) for points (...,3) and quaternions (...,4) with broadcasting
    (quaternion_utils.py:37-54), expanded to  (w^2 - |u|^2) v + 2 (u.v) u + 2 w (u x v)  for q = (u, w).
    For |q| = 1 this is the rotation R(q) v; like the reference it is NOT normalised, so a non-unit
    quaternion also scales by |q|^2 -- which matters for the gradient of the point-constraint loss, taken
    w.r.t. the un-normalised orientation (simple_setup.py:164-175)."""
    u, w = quaternions[..., :3], quaternions[..., 3:]
    return ((w * w - (u * u).sum(-1, keepdim=True)) * points
            + 2.0 * (u * points).sum(-1, keepdim=True) * u
            + 2.0 * w * torch.linalg.cross(u.expand(torch.broadcast_shapes(u.shape, points.shape)),
                                           points.expand(torch.broadcast_shapes(u.shape, points.shape)), dim=-1))


def point_constraint_loss(orientation_q: torch.Tensor, source: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """|| q (x) source (x) conj(q) - target ||  (estimation/losses.py:138-153), batched over (...,4)
    orientations; source, target (3,)."""
    return torch.linalg.norm(quaternion_apply(orientation_q, source) - target, dim=-1)


def to_camera_frames(position: torch.Tensor, orientation: torch.Tensor, camera_positions: torch.Tensor,
                     camera_orientations: torch.Tensor):
    """Object poses (B,3), (B,4) in the world frame -> (V,B,3), (V,B,4) in each of V camera frames
    (simple_setup.py:423-431).  camera_positions (V,3), camera_orientations (V,4) camera-to-world."""
    q_w2c = quaternion_invert(camera_orientations)
    # camera orientations are unit quaternions: the rotation-matrix form of quaternion_apply
    position_c = (rotation_matrix(q_w2c[:, None]) @ (position[None] - camera_positions[:, None])[..., None])[..., 0]
    orientation_c = quaternion_multiply(q_w2c[:, None], orientation[None])
    return position_c, orientation_c


def pull_back(grad_position_c: torch.Tensor, grad_orientation_c: torch.Tensor,
              camera_orientations: torch.Tensor):
    """Adjoint of to_camera_frames: per-view gradients (V,B,3), (V,B,4) w.r.t. the camera-frame poses ->
    gradients (B,3), (B,4) w.r.t. the world-frame pose, summed over the views."""
    q_w2c = quaternion_invert(camera_orientations)
    g_p = (rotation_matrix(q_w2c).transpose(-1, -2)[:, None] @ grad_position_c[..., None])[..., 0].sum(0)
    g_q = (left_matrix(q_w2c).transpose(-1, -2)[:, None] @ grad_orientation_c[..., None])[..., 0].sum(0)
    return g_p, g_q
