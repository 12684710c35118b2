"""numpy restatement of the optimiser side of one render-and-compare iteration -- TEST
INFRASTRUCTURE ONLY (the checker of ``sdfr_hypothesis_step``; nothing under sdfest_b200/ imports it).

Reference, per hypothesis (sdfest/estimation/simple_setup.py):
    :400-406   torch.optim.Adam([position lr 1e-3, orientation 1e-2, scale 1e-3, latent 1e-2])
    :411       norm_orientation = orientation / sqrt(sum(orientation**2))
    :431       render(..., 1 / scale)
    :447-452   loss = depth_weight * loss_depth + pc_weight * loss_pc
    :456-462   backward; optimizer.step(); orientation /= |orientation|
The Adam arithmetic lives in a third-party dependency of the reference, PyTorch
(torch/optim/adam.py ``_single_tensor_adam``, torch 2.11.0 here), defaults betas (0.9, 0.999),
eps 1e-8, no weight decay, no amsgrad:
    m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g^2
    p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
Pinned by tests/test_hypothesis_step.py (CPU) against torch autograd through the same chain
(normalisation, reciprocal, weighted loss) followed by torch.optim.Adam.step() itself.
"""
from __future__ import annotations

import numpy as np


def chain_gradients(orientation, scale, n_overlap, gr_p, gr_q, gr_is, depth_weight,
                    g2_p=None, g2_q=None, g2_s=None):
    """Gradients w.r.t. position, the UN-normalised orientation and scale from the raw render
    gradients (w.r.t. unit quaternion / inv_scale, unnormalised sums over the overlap) and an
    already weighted second set (w.r.t. unit quaternion / scale).  float64."""
    o = np.asarray(orientation, np.float64)
    s = np.asarray(scale, np.float64).reshape(-1)
    n = np.asarray(n_overlap, np.float64).reshape(-1)
    coef = np.where(n > 0, depth_weight / np.where(n > 0, n, 1.0), 0.0)
    z = lambda a, like: np.zeros_like(like, np.float64) if a is None else np.asarray(a, np.float64)  # noqa: E731
    g_p = coef[:, None] * z(gr_p, np.zeros((len(s), 3))) + z(g2_p, np.zeros((len(s), 3)))
    g_u = coef[:, None] * z(gr_q, o) + z(g2_q, o)
    nrm = np.linalg.norm(o, axis=1, keepdims=True)
    q = o / nrm
    g_o = (g_u - q * np.sum(q * g_u, axis=1, keepdims=True)) / nrm
    g_s = -coef * z(gr_is, s).reshape(-1) / (s * s) + z(g2_s, s).reshape(-1)
    return g_p, g_o, g_s


def adam(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One Adam update of torch/optim/adam.py; returns (p, m, v).  t = step count AFTER increment."""
    m = m + (g - m) * (1.0 - beta1)
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    p = p - (lr / bc1) * (m / (np.sqrt(v) / np.sqrt(bc2) + eps))
    return p, m, v


def hypothesis_step(state, loss_sum, n_overlap, gr_p, gr_q, gr_is, depth_weight, point_sum=None,
                    point_weight=0.0, g2_p=None, g2_q=None, g2_s=None, g_latent=None,
                    lrs=(1e-3, 1e-2, 1e-3, 1e-2), betas=(0.9, 0.999), eps=1e-8):
    """state: dict(position (B,3), orientation (B,4), scale (B,), latent (B,L)|None, m, v (B,8+L),
    t int).  Returns (new state, unit_orientation, inv_scale, loss); float64 throughout."""
    pos = np.asarray(state["position"], np.float64)
    ori = np.asarray(state["orientation"], np.float64)
    scale = np.asarray(state["scale"], np.float64).reshape(-1)
    lat = None if state.get("latent") is None else np.asarray(state["latent"], np.float64)
    m, v, t = np.array(state["m"], np.float64), np.array(state["v"], np.float64), int(state["t"]) + 1
    n = np.asarray(n_overlap, np.float64).reshape(-1)
    loss = np.zeros_like(scale)
    if loss_sum is not None:
        # no overlap: NaN, the reference's mean over an empty selection (simple_setup.py:131)
        loss = depth_weight * np.where(n > 0, np.asarray(loss_sum, np.float64) / np.where(n > 0, n, 1), np.nan)
    if point_sum is not None:
        loss = loss + point_weight * np.asarray(point_sum, np.float64)
    g_p, g_o, g_s = chain_gradients(ori, scale, n, gr_p, gr_q, gr_is, depth_weight, g2_p, g2_q, g2_s)
    pos, m[:, 0:3], v[:, 0:3] = adam(pos, g_p, m[:, 0:3], v[:, 0:3], t, lrs[0], *betas, eps)
    ori, m[:, 3:7], v[:, 3:7] = adam(ori, g_o, m[:, 3:7], v[:, 3:7], t, lrs[1], *betas, eps)
    sc, m[:, 7], v[:, 7] = adam(scale, g_s, m[:, 7], v[:, 7], t, lrs[2], *betas, eps)
    if lat is not None and g_latent is not None:
        lat, m[:, 8:], v[:, 8:] = adam(lat, np.asarray(g_latent, np.float64), m[:, 8:], v[:, 8:], t,
                                       lrs[3], *betas, eps)
    ori = ori / np.linalg.norm(ori, axis=1, keepdims=True)  # simple_setup.py:462
    new = dict(position=pos, orientation=ori, scale=sc, latent=lat, m=m, v=v, t=t)
    unit = ori / np.linalg.norm(ori, axis=1, keepdims=True)
    return new, unit, 1.0 / sc, loss


def inlier_counts(depth_obs, depth_est, rel_threshold):
    """(n_inlier, n_valid) of simple_setup.py:177-188 for one image pair, in the float32 arithmetic
    the reference runs in:  rel = |obs - est| / obs;  inliers = count(rel < threshold);  valid =
    count(obs != 0).  A pixel with obs == 0 divides to inf or nan and never counts."""
    obs = np.asarray(depth_obs, np.float32)
    est = np.asarray(depth_est, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(obs - est) / obs
        n_inlier = int(np.count_nonzero(rel < np.float32(rel_threshold)))
    return n_inlier, int(np.count_nonzero(obs))


class BestEstimate:
    """simple_setup.py:190-211 as the code intends it (copies, not the references to the live tensors
    the reference stores at :207-210): keep the parameters with the highest inlier ratio so far."""

    def __init__(self):
        self.ratio, self.iteration, self.params = None, -1, None

    def update(self, n_inlier, n_valid, iteration, params):
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.float32(n_inlier) / np.float32(n_valid)
        if self.ratio is None or ratio > self.ratio:  # :205
            self.ratio, self.iteration = ratio, iteration
            self.params = tuple(None if p is None else np.array(p, copy=True) for p in params)
        return ratio
