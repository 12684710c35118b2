"""Runtime breakdown of the render-and-compare loop.

The reference's protocol (sdfest/estimation/scripts/real_data.py:217-319 with
estimation/configs/runtime_analysis.yaml) wraps the phases of ``SDFPipeline.__call__`` in
``torch.cuda.synchronize()`` + ``time.time()`` pairs and writes a YAML of per-phase seconds.  The
batched loop here has no host synchronisation to hang such timers on, so phases are bracketed with
CUDA events on the loop's stream instead (no extra syncs inside an iteration) and reported in the
same spirit: milliseconds per iteration for decode / render+compare / point loss / backward /
optimiser, for B hypotheses at once, plus the whole iteration replayed as a CUDA graph.
"""
from __future__ import annotations

from typing import Dict

import torch

from .hypotheses import HypothesisOptimizer

PHASES = ("decode", "render_compare", "point_loss", "backward", "optimizer")


def _ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def phase_breakdown(opt: HypothesisOptimizer, iterations: int = 10, warmup: int = 3) -> Dict[str, float]:
    """Average milliseconds per phase of the COMPOSED path (separate operators, so that phases can
    be told apart), B hypotheses per iteration."""
    from ..differentiable_renderer import render_and_compare
    from . import losses

    if opt.optimizer is None:
        raise ValueError("phase_breakdown times the composed path: build the HypothesisOptimizer with "
                         "optimizer='torch' (the fused iteration has no phase boundaries to time)")
    acc = dict.fromkeys(PHASES, 0.0)
    for it in range(warmup + iterations):
        opt.optimizer.zero_grad(set_to_none=True)
        e0 = _ev()
        q = opt.orientation / torch.linalg.norm(opt.orientation, dim=1, keepdim=True)
        grids = opt._grids()
        e1 = _ev()
        loss_d, _, _ = render_and_compare(grids, opt.position, q.contiguous(),
                                          (1.0 / opt.scale).contiguous(), opt.depth_obs,
                                          opt.threshold, opt.camera)
        loss = opt.depth_weight * torch.nan_to_num(loss_d, nan=0.0)
        e2 = _ev()
        if opt.pc_weight and opt.points.shape[0] > 0:
            loss = loss + opt.pc_weight * losses.point_loss(opt.points, opt.position, q, opt.scale, grids)
        e3 = _ev()
        loss.sum().backward()
        e4 = _ev()
        opt.optimizer.step()
        with torch.no_grad():
            opt.orientation /= torch.linalg.norm(opt.orientation, dim=1, keepdim=True)
        e5 = _ev()
        torch.cuda.synchronize()
        if it >= warmup:
            for name, a, b in zip(PHASES, (e0, e1, e2, e3, e4), (e1, e2, e3, e4, e5)):
                acc[name] += a.elapsed_time(b) / iterations
    acc["total"] = sum(acc[p] for p in PHASES)
    return acc


def iteration_ms(opt: HypothesisOptimizer, iterations: int = 30, warmup: int = 3, graph: bool = True) -> float:
    """Milliseconds per iteration of ``opt.step()`` (the product path), optionally graph-replayed."""
    if graph and opt._graph is None:
        opt.capture(warmup=warmup)
    for _ in range(warmup):
        opt.step()
    torch.cuda.synchronize()
    a = _ev()
    for _ in range(iterations):
        opt.step()
    b = _ev()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iterations


# phases of this loop under the names of the reference's timing decorators (real_data.py:228-243):
#   "init"     pipeline._nn_init                 -> the initialisation network call (caller-measured)
#   "decode"   pipeline.vae.decode               -> decode
#   "render"   pipeline.render                   -> render_compare (render + masked L1 are one launch here)
#   "losses"   pipeline._compute_view_losses     -> point_loss (the depth L1 already happened in "render")
#   "backward" pipeline._compute_gradients       -> backward
# plus "optimizer", which the reference does not time separately.
REFERENCE_KEYS = {"decode": "decode", "render_compare": "render", "point_loss": "losses", "backward": "backward",
                  "optimizer": "optimizer"}


def reference_overview(with_decode: Dict[str, float], without_decode: Dict[str, float], iterations_per_run: int,
                       runs: int = 1, init_ms: float = None, config: Dict = None) -> Dict:
    """The YAML document of the reference's ``generate_runtime_overview`` (real_data.py:286-319) from two
    ``phase_breakdown`` results (shape optimisation on / off): ``results_with_decode`` /
    ``results_without_decode``, per phase ``total``, ``total_calls``, ``mean``, ``calls_per_run``,
    ``total_per_run`` -- in SECONDS like the reference, for ``runs`` runs of ``iterations_per_run`` iterations."""
    def stats(breakdown):
        out = {}
        for ours, theirs in REFERENCE_KEYS.items():
            if ours not in breakdown or (theirs == "decode" and breakdown is without_decode):
                continue
            mean = breakdown[ours] * 1e-3
            calls = iterations_per_run * runs
            out[theirs] = {"total": mean * calls, "total_calls": calls, "mean": mean,
                           "calls_per_run": float(iterations_per_run), "total_per_run": mean * iterations_per_run}
        if init_ms is not None:
            out["init"] = {"total": init_ms * 1e-3 * runs, "total_calls": runs, "mean": init_ms * 1e-3,
                           "calls_per_run": 1.0, "total_per_run": init_ms * 1e-3}
        return out

    return {**(config or {}), "results_with_decode": stats(with_decode), "results_without_decode": stats(without_decode)}


def write_yaml(path: str, results: Dict) -> None:
    import yaml

    with open(path, "w") as f:
        yaml.safe_dump(results, f, sort_keys=False)
