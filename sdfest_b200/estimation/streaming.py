"""Render-and-compare for hypotheses that live in HOST memory, streamed through the GPU.

The reference keeps one hypothesis on the device and calls the renderer once per iteration
(estimation/simple_setup.py:432-456).  When SDF grids arrive from the host every step (64 MiB for
64 hypotheses at 64^3) the PCIe copy, not the render, is the long pole: 1.2 ms against 0.3 ms on a
B200.  ``StreamedRenderCompare`` therefore cuts the batch into chunks and runs three queues --
host->device copy of chunk k+1, fused render + masked-L1 compare + backward of chunk k
(``sdfr_compare_fused`` / ``sdfr_scale_grads`` through the C ABI), device->host copy of the small
per-hypothesis results -- so that a step costs max(copy, compute) instead of their sum.  The whole
step can be captured once in a CUDA graph (the library is capture-safe: it never allocates or
synchronises) and replayed with one launch.

Results copied back per hypothesis: loss, overlap count, and the gradients of the loss w.r.t.
position (3), orientation (4) and inverse scale (1) -- 10 floats.  The SDF-grid gradients stay on
the device (``g_sdf``), where the decoder's backward consumes them; ``sdf_grads_to_host=True``
streams them back too.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from .. import _lib
from ..differentiable_renderer import Camera
from ..differentiable_renderer.sdf_renderer import _camera_params


class StreamedRenderCompare:
    def __init__(self, camera: Camera, threshold: float, batch: int, resolution: int, device,
                 chunk: int = 16, sdf_grads_to_host: bool = False, sdf_grad_mode: str = "reference",
                 skew: bool = False):
        self.device = torch.device(device)
        self.B, self.R, self.chunk = int(batch), int(resolution), max(1, min(int(chunk), int(batch)))
        self.W, self.H, self.cx, self.cy, self.fx, self.fy = _camera_params(camera)
        self.threshold = float(threshold)
        self.lib = _lib.lib()
        self.flags = _lib.GRAD_ALL | _lib.ZERO_GRADS | (_lib.SDF_GRAD_EXACT if sdf_grad_mode == "exact" else 0)
        self.sdf_grads_to_host = sdf_grads_to_host
        B, R, dev = self.B, self.R, self.device
        self.RRR = R * R * R
        with torch.cuda.device(dev):
            self.d_sdf = torch.empty(B, self.RRR, device=dev)
            self.d_pos = torch.empty(B, 3, device=dev)
            self.d_quat = torch.empty(B, 4, device=dev)
            self.d_inv = torch.empty(B, device=dev)
            self.d_obs = torch.empty(self.H, self.W, device=dev)
            self.depth = torch.empty(B, self.H, self.W, device=dev)
            # rows of `small`: loss_sum/loss, n_overlap, g_pos(3), g_quat(4), g_inv(1) -- kept as
            # separate contiguous device arrays because the C ABI wants [B,3], [B,4], [B]
            self.sums = torch.empty(2, B, device=dev)
            self.g_sdf = torch.empty(B, self.RRR, device=dev)
            self.g_pos = torch.empty(B, 3, device=dev)
            self.g_quat = torch.empty(B, 4, device=dev)
            self.g_inv = torch.empty(B, device=dev)
            self.d_small = torch.empty(B, 10, device=dev)
            self.h_small = torch.empty(B, 10, pin_memory=True)
            self.h_g_sdf = torch.empty(B, self.RRR, pin_memory=True) if sdf_grads_to_host else None
            self.skewed = None
            if skew:
                n = ctypes.c_longlong(0)
                _lib.check(self.lib.sdfr_skewed_pitches(R, None, None, ctypes.byref(n)), "sdfr_skewed_pitches")
                self.SK = int(n.value)
                self.skewed = torch.empty(B, self.SK, device=dev)
            self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self._graph = None
        self._graph_key = None
        self.h2d_bytes = 4 * (B * self.RRR + B * 8 + self.H * self.W)
        self.d2h_bytes = 4 * (B * 10 + (B * self.RRR if sdf_grads_to_host else 0))

    # ------------------------------------------------------------------------------------------
    def _enqueue(self, h_sdf, h_pos, h_quat, h_inv, h_obs, main):
        """Enqueue one step; `main` is the stream the caller's work is ordered on."""
        lib, B, R, RRR = self.lib, self.B, self.R, self.RRR
        s_in, s_cmp, s_out = self.s_in, self.s_cmp, self.s_out
        s_in.wait_stream(main)
        s_cmp.wait_stream(main)
        s_out.wait_stream(main)
        with torch.cuda.stream(s_in):
            self.d_pos.copy_(h_pos, non_blocking=True)
            self.d_quat.copy_(h_quat, non_blocking=True)
            self.d_inv.copy_(h_inv, non_blocking=True)
            self.d_obs.copy_(h_obs, non_blocking=True)
        h_sdf = h_sdf.view(B, RRR)
        f4 = 4
        for b0 in range(0, B, self.chunk):
            n = min(self.chunk, B - b0)
            with torch.cuda.stream(s_in):
                self.d_sdf[b0:b0 + n].copy_(h_sdf[b0:b0 + n], non_blocking=True)
            s_cmp.wait_stream(s_in)
            with torch.cuda.stream(s_cmp):
                st = s_cmp.cuda_stream
                if self.skewed is not None:
                    _lib.check(lib.sdfr_skew_grids(self.d_sdf.data_ptr() + b0 * RRR * f4, R, RRR, n,
                                                   self.skewed.data_ptr() + b0 * self.SK * f4, self.SK, st),
                               "sdfr_skew_grids")
                    src, stride, layout = self.skewed.data_ptr() + b0 * self.SK * f4, self.SK, _lib.LAYOUT_SKEWED
                else:
                    src, stride, layout = self.d_sdf.data_ptr() + b0 * RRR * f4, RRR, _lib.LAYOUT_DENSE
                grads = (self.g_sdf.data_ptr() + b0 * RRR * f4, RRR, self.g_pos.data_ptr() + b0 * 3 * f4,
                         self.g_quat.data_ptr() + b0 * 4 * f4, self.g_inv.data_ptr() + b0 * f4)
                _lib.check(lib.sdfr_compare_fused(
                    src, R, stride, layout, self.d_pos.data_ptr() + b0 * 3 * f4,
                    self.d_quat.data_ptr() + b0 * 4 * f4, self.d_inv.data_ptr() + b0 * f4, n,
                    self.W, self.H, self.cx, self.cy, self.fx, self.fy, self.threshold,
                    self.d_obs.data_ptr(), 0, self.depth.data_ptr() + b0 * self.H * self.W * f4,
                    self.sums[0].data_ptr() + b0 * f4, self.sums[1].data_ptr() + b0 * f4, *grads,
                    self.flags, None, st), "sdfr_compare_fused")
                _lib.check(lib.sdfr_scale_grads(
                    self.sums[1].data_ptr() + b0 * f4, None, R, n, *grads, _lib.GRAD_ALL, None, 0, st),
                    "sdfr_scale_grads")
            if self.sdf_grads_to_host:
                s_out.wait_stream(s_cmp)
                with torch.cuda.stream(s_out):
                    self.h_g_sdf[b0:b0 + n].copy_(self.g_sdf[b0:b0 + n], non_blocking=True)
        with torch.cuda.stream(s_cmp):
            torch.div(self.sums[0], self.sums[1], out=self.d_small[:, 0])
            self.d_small[:, 1] = self.sums[1]
            self.d_small[:, 2:5] = self.g_pos
            self.d_small[:, 5:9] = self.g_quat
            self.d_small[:, 9] = self.g_inv
        s_out.wait_stream(s_cmp)
        with torch.cuda.stream(s_out):
            self.h_small.copy_(self.d_small, non_blocking=True)
        main.wait_stream(s_out)
        main.wait_stream(s_cmp)
        main.wait_stream(s_in)

    def __call__(self, h_sdf: torch.Tensor, h_position: torch.Tensor, h_orientation: torch.Tensor,
                 h_inv_scale: torch.Tensor, h_depth_obs: torch.Tensor, graph: bool = False,
                 sync: bool = True) -> Dict[str, Optional[torch.Tensor]]:
        """One step from host buffers (pinned for the copies to overlap).  With ``graph=True`` the
        step is captured once per set of host buffers and replayed afterwards.  Returns host views
        (valid after the call when ``sync`` is true) and the device-resident SDF gradients."""
        args = (h_sdf, h_position, h_orientation, h_inv_scale, h_depth_obs)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            if graph:
                key = tuple(t.data_ptr() for t in args)
                if self._graph is None or self._graph_key != key:
                    self._enqueue(*args, main)  # warm-up outside capture
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    cap = torch.cuda.Stream(self.device)
                    with torch.cuda.graph(g, stream=cap):
                        self._enqueue(*args, torch.cuda.current_stream(self.device))
                    self._graph, self._graph_key = g, key
                self._graph.replay()
            else:
                self._enqueue(*args, main)
            if sync:
                main.synchronize()
        hs = self.h_small
        return {"loss": hs[:, 0], "n_overlap": hs[:, 1], "g_position": hs[:, 2:5],
                "g_orientation": hs[:, 5:9], "g_inv_scale": hs[:, 9], "g_sdf_device": self.g_sdf.view(
                    self.B, self.R, self.R, self.R), "g_sdf_host": self.h_g_sdf, "depth_device": self.depth}


class StreamedDecodeRenderCompare:
    """Render-and-compare for hypotheses whose LATENTS live in host memory: what the reference's callers
    actually hold (estimation/simple_setup.py:414 decodes the grid on the device from 8 floats).

    Per step the host hands over, from pinned memory, latent (B,L), position (B,3), unit orientation (B,4),
    scale (B,) and the observed depth map (H,W) -- 1.2 MB for 64 hypotheses at 640x480 instead of the 68 MB of
    grids ``StreamedRenderCompare`` ships -- and reads back per hypothesis the masked-L1 loss, the overlap
    count and the gradients of the loss w.r.t. latent, position, orientation and scale.  On the device:
    decoder trunk + ``sdfr_decoder_tail_forward`` (grids written once, in the skewed layout) ->
    ``sdfr_compare_fused`` -> tail adjoint -> trunk backward (``decode_render_compare``).  The whole step --
    copies included -- is captured once per set of host buffers in a CUDA graph and replayed.
    """

    def __init__(self, decoder, camera: Camera, threshold: float, batch: int, latent_size: int, device,
                 depth_weight: float = 1.0):
        from .decoder import FusedTailDecoder

        if not isinstance(decoder, FusedTailDecoder):
            raise TypeError("decoder must be a FusedTailDecoder")
        self.device = torch.device(device)
        self.decoder, self.camera, self.threshold = decoder, camera, float(threshold)
        self.B, self.L, self.depth_weight = int(batch), int(latent_size), float(depth_weight)
        self.W, self.H = int(camera.width), int(camera.height)
        B, L, dev = self.B, self.L, self.device
        with torch.cuda.device(dev):
            self.d_latent = torch.empty(B, L, device=dev)
            self.d_pos = torch.empty(B, 3, device=dev)
            self.d_quat = torch.empty(B, 4, device=dev)
            self.d_scale = torch.empty(B, device=dev)
            self.d_obs = torch.empty(self.H, self.W, device=dev)
            self.d_out = torch.empty(B, 10 + L, device=dev)
            self.h_out = torch.empty(B, 10 + L, pin_memory=True)
        self.depth = None  # the rendered depth maps of the last step (device)
        self._graph = None
        self._graph_key = None
        self.h2d_bytes = 4 * (B * (L + 8) + self.H * self.W)
        self.d2h_bytes = 4 * B * (10 + L)

    def _enqueue(self, h_latent, h_pos, h_quat, h_scale, h_obs):
        from .fused import decode_render_compare

        for d, h in ((self.d_latent, h_latent), (self.d_pos, h_pos), (self.d_quat, h_quat),
                     (self.d_scale, h_scale), (self.d_obs, h_obs)):
            d.copy_(h.view(d.shape), non_blocking=True)
        leaves = [t.detach().requires_grad_(True) for t in (self.d_latent, self.d_pos, self.d_quat, self.d_scale)]
        lat, p, q, s = leaves
        w, b = self.decoder.tail_parameters()
        loss, depth, n, _ = decode_render_compare(
            self.decoder.trunk(lat), w, b, p, q, s, self.d_obs, None, self.decoder.volume_size, self.threshold,
            self.camera, base=self.decoder.base, depth_weight=self.depth_weight, pc_weight=0.0)
        g_lat, g_p, g_q, g_s = torch.autograd.grad(loss.sum(), leaves)
        self.depth = depth
        torch.cat([loss.detach()[:, None], n[:, None], g_p, g_q, g_s[:, None], g_lat], 1, out=self.d_out)
        self.h_out.copy_(self.d_out, non_blocking=True)

    def __call__(self, h_latent: torch.Tensor, h_position: torch.Tensor, h_orientation: torch.Tensor,
                 h_scale: torch.Tensor, h_depth_obs: torch.Tensor, graph: bool = True,
                 sync: bool = True) -> Dict[str, Optional[torch.Tensor]]:
        """One step from (pinned) host buffers; returns host views, valid after the call when ``sync``."""
        args = (h_latent, h_position, h_orientation, h_scale, h_depth_obs)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            if graph:
                key = tuple(t.data_ptr() for t in args)
                if self._graph is None or self._graph_key != key:
                    side = torch.cuda.Stream(self.device)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):  # warm-up outside capture (cuBLAS workspaces, autograd)
                        for _ in range(2):
                            self._enqueue(*args)
                    main.wait_stream(side)
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._enqueue(*args)
                    self._graph, self._graph_key = g, key
                self._graph.replay()
            else:
                self._enqueue(*args)
            if sync:
                main.synchronize()
        hs, L = self.h_out, self.L
        return {"loss": hs[:, 0], "n_overlap": hs[:, 1], "g_position": hs[:, 2:5], "g_orientation": hs[:, 5:9],
                "g_scale": hs[:, 9], "g_latent": hs[:, 10:10 + L], "depth_device": self.depth}
