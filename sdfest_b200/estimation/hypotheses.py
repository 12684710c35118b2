"""Batched render-and-compare: the loop body of SDFPipeline.__call__ for B hypotheses at once.

Reference: estimation/simple_setup.py:400-470 optimises ONE hypothesis with a Python loop of
decode -> render -> losses -> backward -> Adam, ~dozens of tiny kernels and two host syncs per
iteration.  Here B pose/scale(/latent) hypotheses advance together: one fused render-and-compare
launch pair (libsdfrender.so), batched torch ops for the caller-side pieces, no host
synchronisation inside the loop, and -- across GPUs -- hypotheses sharded per rank with one tiny
all_gather of the per-hypothesis losses (the reference has no distributed code at all).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from ..differentiable_renderer import Camera, render_and_compare
from . import losses
from .decoder import FusedTailDecoder
from .fused import decode_render_compare


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_total hypotheses owned by `rank` (remainder to low ranks)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(n_total, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def gather_losses(local: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks' per-hypothesis losses, concatenated in rank order (equal shard sizes use
    all_gather_into_tensor -- NCCL over NVLink on GPUs, gloo on CPU in the tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = torch.tensor([local.numel()], device=local.device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    sizes = [int(s.item()) for s in all_sizes]
    if len(set(sizes)) == 1:
        out = local.new_empty(world * local.numel())
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = local.new_zeros(pad)
    buf[: local.numel()] = local
    out = [local.new_empty(pad) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)])


def global_best(local_losses: torch.Tensor, lo: int, group=None):
    """(global index, loss) of the best hypothesis over all ranks; NaN losses never win."""
    all_losses = gather_losses(torch.nan_to_num(local_losses, nan=float("inf")), group)
    idx = int(torch.argmin(all_losses).item())
    return idx, float(all_losses[idx].item())


class HypothesisOptimizer:
    """Adam on B hypotheses of (position, orientation, scale[, latent]) against one observation.

    Learning rates and loss weights default to the reference's (simple_setup.py:400-405,
    estimation/configs/default.yaml:14-16).  ``decoder`` maps latents (B,L) to grids
    (B,1,R,R,R) (e.g. ``SDFVAE.decode``); without it ``sdf`` holds fixed grids (B|1,R,R,R).
    """

    def __init__(self, camera: Camera, threshold: float, depth_obs: torch.Tensor,
                 position: torch.Tensor, orientation: torch.Tensor, scale: torch.Tensor,
                 sdf: Optional[torch.Tensor] = None, latent: Optional[torch.Tensor] = None,
                 decoder: Optional[Callable] = None, depth_weight: float = 1.0,
                 pc_weight: float = 3.0, max_points: int = 0, group=None):
        if (decoder is None) == (sdf is None):
            raise ValueError("give either fixed `sdf` grids or a `decoder` with `latent`")
        self.camera, self.threshold, self.group = camera, float(threshold), group
        self.depth_obs = depth_obs.contiguous()
        self.depth_weight, self.pc_weight = depth_weight, pc_weight
        self.position = position.detach().clone().requires_grad_(True)
        self.orientation = orientation.detach().clone().requires_grad_(True)
        self.scale = scale.detach().clone().requires_grad_(True)
        self.decoder, self.sdf = decoder, sdf
        groups = [{"params": [self.position], "lr": 1e-3}, {"params": [self.orientation], "lr": 1e-2},
                  {"params": [self.scale], "lr": 1e-3}]
        self.latent = None
        if decoder is not None:
            self.latent = latent.detach().clone().requires_grad_(True)
            groups.append({"params": [self.latent], "lr": 1e-2})
        self.optimizer = torch.optim.Adam(groups, capturable=self.position.is_cuda)
        # observed points, once (the only host sync), sub-sampled to a fixed size
        pts = losses.depth_to_pointcloud(self.depth_obs, camera)
        if max_points and pts.shape[0] > max_points:
            sel = torch.randperm(pts.shape[0], device=pts.device,
                                 generator=torch.Generator(pts.device).manual_seed(0))[:max_points]
            pts = pts[sel]
        self.points = pts.contiguous()
        self.last_losses = None
        self._graph = None

    def _grids(self):
        if self.decoder is None:
            return self.sdf
        g = self.decoder(self.latent)
        return g[:, 0].contiguous() if g.dim() == 5 else g.contiguous()

    def capture(self, warmup: int = 3) -> None:
        """Capture one whole iteration (decode, render-and-compare, point loss, backward, Adam,
        renormalisation) in a CUDA graph; step() replays it afterwards.  The reference loop
        cannot be captured: it synchronises with the host twice per iteration
        (simple_setup.py:131, pointset_utils.py:60) and its renderer launches on the legacy
        default stream (sdf_renderer_cuda.cu:495)."""
        if not self.position.is_cuda:
            raise RuntimeError("CUDA graphs need CUDA tensors")
        side = torch.cuda.Stream(self.position.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            self._static_loss = self._eager_step()
        self._graph = graph

    def step(self) -> torch.Tensor:
        """One iteration (simple_setup.py:408-470); returns the detached per-hypothesis loss."""
        if self._graph is not None:
            self._graph.replay()
            self.last_losses = self._static_loss
            return self.last_losses
        return self._eager_step()

    def _fused_loss(self, q: torch.Tensor) -> torch.Tensor:
        """Decoder tail, render-and-compare and point loss chained through the C ABI
        (estimation/fused.py): no dense grid, no separate layout / scaling / add passes."""
        dec = self.decoder
        w, b = dec.tail_parameters()
        loss, _, _, _ = decode_render_compare(
            dec.trunk(self.latent), w, b, self.position, q.contiguous(), self.scale,
            self.depth_obs, self.points if self.pc_weight and self.points.shape[0] > 0 else None,
            dec.volume_size, self.threshold, self.camera, base=dec.base,
            depth_weight=self.depth_weight, pc_weight=self.pc_weight)
        return loss

    def _eager_step(self) -> torch.Tensor:
        self.optimizer.zero_grad(set_to_none=True)
        q = self.orientation / torch.linalg.norm(self.orientation, dim=1, keepdim=True)
        if isinstance(self.decoder, FusedTailDecoder) and self.position.is_cuda:
            loss = self._fused_loss(q)
            loss.sum().backward()
            self.optimizer.step()
            with torch.no_grad():
                self.orientation /= torch.linalg.norm(self.orientation, dim=1, keepdim=True)
            self.last_losses = loss.detach()
            return self.last_losses
        grids = self._grids()
        loss_depth, _, _ = render_and_compare(grids, self.position, q.contiguous(),
                                              (1.0 / self.scale).contiguous(), self.depth_obs,
                                              self.threshold, self.camera)
        loss = self.depth_weight * torch.nan_to_num(loss_depth, nan=0.0)
        if self.pc_weight and self.points.shape[0] > 0:
            # NB: the reference passes the un-normalised quaternion and lets pc_loss normalise it
            # (simple_setup.py:436-443, losses.py:56); q is already unit here, same value
            loss = loss + self.pc_weight * losses.point_loss(
                self.points, self.position, q, self.scale,
                grids if grids.dim() == 4 else grids[None])
        loss.sum().backward()
        self.optimizer.step()
        with torch.no_grad():
            self.orientation /= torch.linalg.norm(self.orientation, dim=1, keepdim=True)
        self.last_losses = loss.detach()
        return self.last_losses

    def run(self, iterations: int, gather_every: int = 0):
        """`iterations` steps; with gather_every=k the losses of all ranks are all-gathered every
        k-th iteration (and always after the last one).  Returns the gathered losses."""
        out = None
        for it in range(1, iterations + 1):
            local = self.step()
            if (gather_every and it % gather_every == 0) or it == iterations:
                out = gather_losses(local, self.group)
        return out
