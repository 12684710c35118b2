"""Golden vectors for the loop's result selection, from the REFERENCE's own methods.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden_selection.py

``sdfest/estimation/simple_setup.py`` cannot be imported (open3d, yoco, a JIT CUDA build), but the two
methods on this path -- ``SDFPipeline._compute_inlier_ratio`` (:177-188) and ``_update_best_estimate``
(:190-211) -- only use torch: their source is cut out of the file with ``ast`` and executed UNCHANGED
against a small stand-in for ``self`` (the three attributes they touch).  Stored in
``selection.npz``: seeded observed / estimated depth images with every special case (no observation
with and without an estimate, missed pixels, errors around the threshold), the threshold, the inlier
ratio the reference returns for each pair, and -- driving ``_update_best_estimate`` over the sequence
with parameters that change every iteration -- the index it reports as best, together with the
parameters it then hands back, which are the LAST ones: the method stores references to the live
tensors (:207-210) that the loop keeps updating in place.
"""
from __future__ import annotations

import ast
import os
import textwrap

import numpy as np
import torch

REF = "/root/reference/sdfest/estimation/simple_setup.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_methods():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SDFPipeline")
    ns = {"torch": torch}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("_compute_inlier_ratio", "_update_best_estimate"):
            exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns["_compute_inlier_ratio"], ns["_update_best_estimate"]


class Stand:
    """What the two methods read and write on ``self``."""

    def __init__(self, threshold, compute):
        self._relative_inlier_threshold = threshold
        self._best_inlier_ratio = None
        self._compute = compute

    def _compute_inlier_ratio(self, depth_input, depth_estimate):
        return self._compute(self, depth_input, depth_estimate)


def main():
    compute, update = reference_methods()
    g = np.random.default_rng(42)
    n, H, W, thr = 6, 24, 32, 0.03
    obs = (0.4 + 0.3 * g.random((H, W))).astype(np.float32)
    obs[g.random((H, W)) < 0.3] = 0.0
    est = (obs[None] * (1 + 0.08 * (g.random((n, H, W)) - 0.5))).astype(np.float32)
    est[g.random((n, H, W)) < 0.2] = 0.0
    est[(obs[None] == 0) & (g.random((n, H, W)) < 0.5)] = 0.5
    est[3] = obs * np.float32(1.01)  # the best iteration, in the middle of the sequence
    stand = Stand(thr, compute)
    position, scale = torch.zeros(3), torch.ones(1)
    ratios, best_so_far = [], []
    for it in range(n):
        position += 0.01  # in place, as optimizer.step() updates the live tensors
        scale *= 1.01
        r = update(stand, torch.tensor(obs), torch.tensor(est[it]), position, None, scale, None)
        ratios.append(float(r))
        best_so_far.append(float(stand._best_inlier_ratio))
    np.savez_compressed(
        os.path.join(HERE, "selection.npz"), obs=obs, est=est, threshold=np.float32(thr),
        ratios=np.array(ratios, np.float32), best_so_far=np.array(best_so_far, np.float32),
        returned_position=stand._best_position.numpy().copy(), last_position=position.numpy().copy(),
        positions=np.stack([np.full(3, 0.01 * (k + 1), np.float32) for k in range(n)]))
    print("ratios", ratios, "best", best_so_far[-1],
          "returned == last iterate:", bool(torch.equal(stand._best_position, position)))


if __name__ == "__main__":
    main()
