"""BASELINE configs 1, 3 and 5, measured on the GPU box.

    python scripts/gpu_configs.py [tag]                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \\
        --master-port 29544 scripts/gpu_configs.py [tag]                    # G GPUs (C5 shards)

C1  random-init SDFVAE decoder (latent 8 -> 64^3, the reference's architecture) + ONE 640x480 depth render,
    forward and backward down to the latent, through the public autograd API (render_depth_gpu).  Three
    arms on the same decoder weights: (a) the plain PyTorch decoder + this library's renderer -- what a user
    of the reference gets by switching the renderer alone; (b) the fused-tail decoder + renderer
    (FusedTailDecoder); (c) the plain PyTorch decoder + the reference's own CUDA extension behind a
    minimal autograd wrapper (sdf_renderer.py:296-357 needs open3d at import), when oracle/_ref exists.
    The decoder decodes residuals around an analytic mug (a random-init decoder has no surface).  Rank 0.
C3  multi-instance frame: 16 objects with independent 128^3 grids on a 4x4 lattice rendered into ONE
    1280x720 depth map (per-pixel minimum positive depth), forward + backward with all four gradients.
    Ours: sdfr_forward_composite + sdfr_backward_composite (2 launches).  Beside it, when oracle/_ref is
    present, the reference's own CUDA extension driven the only way it can be: 16 forward calls, a
    torch min-composite, 16 backward calls with the winner-masked upstream image.  Rank 0 only.
C5  end-to-end analysis-by-synthesis: 256 object instances sharded over the ranks (32 per GPU at 8
    GPUs; weak scaling keeps 32 per GPU), each with its own Redwood-shaped observation (640x480,
    fx=fy=525, cx=319.5, cy=239.5, pixel centre 0; estimation/configs/redwood.yaml:4-12).  Per batch:
    a random-init initialisation network of the reference's architecture (PointNet 3->128x4->1024,
    dense + residual + batch norm, pose head 1024->512->256->128; estimation/configs/models/mug.yaml:
    97-111) on PyTorch -- not the hot path, the north star keeps it there -- then 100 iterations of the
    fused render-and-compare loop with the decoder inside (HypothesisOptimizer with one observation
    per instance, CUDA-graph replay) and an all_gather of the per-instance losses every iteration.
    A random-init network does not emit poses, so its raw outputs are squashed into a plausible range
    around the cloud centroid (documented below); the work per iteration does not depend on that.

Writes gpurun_out/<tag>_configs_n<G>.json and prints it.  Time = CUDA events, max over ranks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200 import synthetic as syn  # noqa: E402
from sdfest_b200.differentiable_renderer import (Camera, render_depth_batched,  # noqa: E402
                                                 render_depth_composite)
from sdfest_b200.estimation import HypothesisOptimizer, depth_to_pointclouds  # noqa: E402
from sdfest_b200.estimation.hypotheses import gather_losses  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
THR = 0.005
out = {"n_gpus": world}


def timed(fn, n, warm=3, flush=None):
    """Mean ms per call over n calls, CUDA events around each call (L2 flushed before it if asked)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / n


# ---------------------------------------------------------------------------------------------
# C1
# ---------------------------------------------------------------------------------------------
def config1():
    from sdfest_b200.differentiable_renderer import render_depth_gpu

    W, H, R = 640, 480, 64
    cam = Camera(W, H, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
    fused = syn.residual_decoder(R, dev, syn.sdf_mug(R, dev))
    plain, base = fused.decoder, fused.base  # the same weights as a plain torch module (cuDNN)
    hyp = syn.make_hypotheses(1, seed=0, device=dev)
    z0 = 0.3 * torch.randn(1, 8, generator=torch.Generator().manual_seed(1)).to(dev)
    up = torch.randn(H, W, generator=torch.Generator().manual_seed(2)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def leaves():
        return [t.detach().clone().requires_grad_(True)
                for t in (z0, hyp["position"][0], hyp["orientation"][0], hyp["inv_scale"])]

    def run(decode, render):
        z, p, q, s = leaves()
        d = render(decode(z), p, q, s)
        d.backward(up)
        return d.detach(), z.grad, p.grad, q.grad, s.grad

    dec_plain = lambda z: (plain(z)[0, 0] + base).contiguous()  # noqa: E731
    dec_fused = lambda z: fused(z)[0, 0]  # noqa: E731
    ours = lambda g, p, q, s: render_depth_gpu(g, p, q, s, threshold=THR, camera=cam)  # noqa: E731
    res = {"workload": "C1: random-init SDFVAE decoder (latent 8 -> 64^3) + one 640x480 depth render, forward + "
                       "backward to latent / pose / scale, L2 flushed before every step"}
    ref_out = run(dec_plain, ours)
    res["hit_pixels"] = int((ref_out[0] > 0).sum())
    ms_a = timed(lambda: run(dec_plain, ours), 30, flush=flush)
    ms_b = timed(lambda: run(dec_fused, ours), 30, flush=flush)
    res["torch_decoder_our_renderer"] = {"ms": ms_a, "mpix_per_s": W * H / (ms_a * 1e-3) / 1e6}
    res["fused_tail_decoder_our_renderer"] = {"ms": ms_b, "mpix_per_s": W * H / (ms_b * 1e-3) / 1e6}
    ms_dec = timed(lambda: dec_plain(z0.clone().requires_grad_(True)).sum().backward(), 30, flush=flush)
    res["torch_decoder_alone_fwd_bwd_ms"] = ms_dec
    # (d) the same step (b) captured ONCE as a CUDA graph and replayed: what the step costs on the GPU when the
    # host no longer issues ~60 launches per call.  The library's operators launch on the caller's stream and never
    # allocate or synchronise, so the whole autograd step is capturable; the reference's extension launches on the
    # legacy default stream (sdf_renderer_cuda.cu:497, 538) and cannot be captured.
    try:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                run(dec_fused, ours)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g_out = run(dec_fused, ours)
        graph.replay()
        torch.cuda.synchronize()
        eager = run(dec_fused, ours)
        ms_d = timed(graph.replay, 30, flush=flush)
        res["fused_tail_decoder_our_renderer_cuda_graph"] = {
            "ms": ms_d, "mpix_per_s": W * H / (ms_d * 1e-3) / 1e6,
            "max_abs_diff_vs_eager": [float((a - b).abs().max()) for a, b in zip(g_out, eager)]}
    except Exception as e:  # noqa: BLE001
        res["fused_tail_decoder_our_renderer_cuda_graph"] = {"unavailable": str(e)[:300]}
    try:
        from oracle import build_ref
        ext = build_ref.load_module()
    except Exception as e:  # noqa: BLE001
        ext, res["reference_cuda_ext"] = None, {"unavailable": str(e)[:200]}
    if ext is not None:
        class RefRender(torch.autograd.Function):  # sdf_renderer.py:296-357, minus the open3d import
            @staticmethod
            def forward(ctx, sdf, p, q, s):
                (d,) = ext.forward(sdf, p, q, s, W, H, 320.0, 240.0, 320.0, 320.0, THR)
                ctx.save_for_backward(d, sdf, p, q, s)
                return d

            @staticmethod
            def backward(ctx, g):
                d, sdf, p, q, s = ctx.saved_tensors
                return tuple(ext.backward(g.contiguous(), d, sdf, p, q, s, W, H, 320.0, 240.0, 320.0, 320.0))

        theirs = run(dec_plain, RefRender.apply)
        ms_c = timed(lambda: run(dec_plain, RefRender.apply), 30, flush=flush)
        both = (theirs[0] > 0) & (ref_out[0] > 0)
        rel = ((theirs[0] - ref_out[0]).abs() / theirs[0].clamp(min=1e-6))[both]

        def gerr(a, b):
            return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))

        res["reference_cuda_ext"] = {
            "ms": ms_c, "mpix_per_s": W * H / (ms_c * 1e-3) / 1e6, "speedup_renderer_swapped": ms_c / ms_a,
            "speedup_fused_decoder_too": ms_c / ms_b,
            "speedup_cuda_graph": (ms_c / res["fused_tail_decoder_our_renderer_cuda_graph"]["ms"]
                                   if "ms" in res["fused_tail_decoder_our_renderer_cuda_graph"] else None),
            "hit_mask_agreement": float(((theirs[0] > 0) == (ref_out[0] > 0)).float().mean()),
            "depth_within_1e-5_rel": float((rel <= 1e-5).float().mean()),
            "latent_grad_max_norm_rel_err": gerr(ref_out[1], theirs[1]),
            "position_grad_max_norm_rel_err": gerr(ref_out[2], theirs[2]),
            "orientation_grad_max_norm_rel_err": gerr(ref_out[3], theirs[3]),
            "inv_scale_grad_max_norm_rel_err": gerr(ref_out[4], theirs[4])}
    return res


# ---------------------------------------------------------------------------------------------
# C3
# ---------------------------------------------------------------------------------------------
def config3():
    K, R, W, H = 16, 128, 1280, 720
    cam = Camera(W, H, 640.0, 640.0, 640.0, 360.0, pixel_center=0.5)
    names = ("mug", "bowl", "bottle")
    grids = torch.stack([syn.category_grid(names[k % 3], R, dev, shape_param=0.3 * ((k % 5) - 2) / 2)
                         for k in range(K)]).contiguous()
    g = torch.Generator().manual_seed(3)
    z = -(0.6 + 0.6 * torch.rand(K, generator=g))
    ix, iy = torch.arange(K) % 4, torch.arange(K) // 4
    pos = torch.stack([(ix - 1.5) * 0.42 * (-z), (iy - 1.5) * 0.24 * (-z), z], 1)  # lattice in the image
    quat = syn.random_unit_quaternions(K, g)
    scale = 0.08 + 0.07 * torch.rand(K, generator=g)
    pos, quat, inv_s = (t.float().contiguous().to(dev) for t in (pos, quat, 1.0 / scale))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    up = torch.randn(H, W, generator=torch.Generator().manual_seed(4)).to(dev)
    leaves = [t.clone().requires_grad_(True) for t in (grids, pos, quat, inv_s)]

    def ours():
        depth, winner = render_depth_composite(*leaves, THR, cam)
        torch.autograd.backward(depth, up)
        for t in leaves:
            t.grad = None
        return depth, winner

    depth, winner = ours()
    ms = timed(ours, 20, flush=flush)
    res = {"workload": "C3: 16 objects x 128^3 grids composited into one 1280x720 depth map, fwd+bwd "
                       "(all four gradients), L2 flushed before every step",
           "ms_per_frame": ms, "mpix_per_s": W * H / (ms * 1e-3) / 1e6,
           "object_mpix_per_s": K * W * H / (ms * 1e-3) / 1e6,
           "hit_fraction": float((winner >= 0).float().mean()),
           "objects_visible": int(torch.unique(winner[winner >= 0]).numel())}
    try:
        from oracle import build_ref
        ext = build_ref.load_module()
    except Exception as e:  # noqa: BLE001
        ext, res["reference_cuda_ext"] = None, {"unavailable": str(e)[:200]}
    if ext is not None:
        # the reference kernel hard-codes R = 64 (cu:225-230, :327) and renders one object per call: it
        # gets the same 16 objects at 64^3 (the only resolution it supports) and a torch min-composite
        g64 = torch.stack([syn.category_grid(names[k % 3], 64, dev, shape_param=0.3 * ((k % 5) - 2) / 2)
                           for k in range(K)]).contiguous()

        def ref():
            ds = []
            for k in range(K):
                (d,) = ext.forward(g64[k], pos[k], quat[k], inv_s[k:k + 1], W, H, 640.0, 360.0, 640.0,
                                   640.0, THR)
                ds.append(d)
            stack = torch.stack(ds)
            far = torch.where(stack > 0, stack, torch.full_like(stack, float("inf")))
            best, win = far.min(0)
            comp = torch.where(torch.isfinite(best), best, torch.zeros_like(best))
            for k in range(K):
                gk = torch.where((win == k) & (comp > 0), up, torch.zeros_like(up))
                ext.backward(gk, ds[k], g64[k], pos[k], quat[k], inv_s[k:k + 1], W, H, 640.0, 360.0,
                             640.0, 640.0)
            return comp

        comp = ref()
        ms_ref = timed(ref, 5, warm=1, flush=flush)
        with torch.no_grad():
            d64, _ = render_depth_composite(g64, pos, quat, inv_s, THR, cam)
        both = (comp > 0) & (d64 > 0)
        rel = ((comp - d64).abs() / comp.clamp(min=1e-6))[both]
        res["reference_cuda_ext"] = {
            "ms_per_frame": ms_ref, "mpix_per_s": W * H / (ms_ref * 1e-3) / 1e6,
            "speedup": ms_ref / ms,
            "what": "the same 16 objects at 64^3 (the reference kernel hard-codes R = 64): 16 x "
                    "sdf_renderer_cpp.forward + torch min-composite + 16 x sdf_renderer_cpp.backward",
            "hit_mask_agreement_at_64": float(((comp > 0) == (d64 > 0)).float().mean()),
            "depth_within_1e-5_rel_at_64": float((rel <= 1e-5).float().mean()) if rel.numel() else None}
    return res


# ---------------------------------------------------------------------------------------------
# C5
# ---------------------------------------------------------------------------------------------
class PointNetBackbone(nn.Module):
    """Shared per-point MLP with batch norm; `dense`: every layer but the last also sees the set-wise
    maximum of its own output; `residual`: skip connection wherever the shapes agree; global max pool
    (the architecture of sdfest/initialization/pointnet.py:7-96 with mug.yaml:99-105, random init)."""

    def __init__(self, widths=(128, 128, 128, 128, 1024)):
        super().__init__()
        ins = [3] + [2 * w for w in widths[:-1]]
        self.fc = nn.ModuleList(nn.Linear(i, o) for i, o in zip(ins, widths))
        self.bn = nn.ModuleList(nn.BatchNorm1d(o) for o in widths)

    def forward(self, x):  # (N, M, 3)
        prev = x
        for i, (fc, bn) in enumerate(zip(self.fc, self.bn)):
            y = torch.relu(bn(fc(prev).flatten(0, 1)).view(x.shape[0], x.shape[1], -1))
            if i + 1 < len(self.fc):
                y = torch.cat([y, y.max(1, keepdim=True)[0].expand_as(y)], 2)
            prev = prev + y if prev.shape == y.shape else y
        return prev.max(1)[0]


class PoseHead(nn.Module):
    """1024 -> 512 -> 256 -> 128 -> latent + position 3 + scale 1 + quaternion 4
    (sdfest/initialization/sdf_pose_network.py:9-121, orientation_repr "quaternion")."""

    def __init__(self, latent=8, widths=(512, 256, 128)):
        super().__init__()
        ins = [1024] + list(widths[:-1])
        self.fc = nn.ModuleList(nn.Linear(i, o) for i, o in zip(ins, widths))
        self.bn = nn.ModuleList(nn.BatchNorm1d(o) for o in widths)
        self.out = nn.Linear(widths[-1], latent + 8)
        self.latent = latent

    def forward(self, f):
        for fc, bn in zip(self.fc, self.bn):
            f = torch.relu(bn(fc(f)))
        o = self.out(f)
        L = self.latent
        return o[:, :L], o[:, L:L + 3], o[:, L + 3], nn.functional.normalize(o[:, L + 4:], dim=1)


def config5():
    TOTAL = int(os.environ.get("C5_TOTAL", str(32 * world)))  # 256 at 8 GPUs; 32 per GPU (weak scaling)
    ITER = int(os.environ.get("C5_ITER", "100"))
    NPTS = 2048  # points per instance fed to the initialisation network
    K = TOTAL // world
    W, H, R = 640, 480, 64
    cam = Camera(W, H, 525.0, 525.0, 319.5, 239.5, pixel_center=0.0)
    torch.manual_seed(100 + rank)
    truth = syn.make_hypotheses(K, seed=100 + rank, device=dev, base_position=(0.0, 0.0, -0.7),
                                pos_sigma=0.05, rot_deg=180.0, scale_rel=0.2)
    dec = syn.residual_decoder(R, dev, syn.sdf_mug(R, dev))
    for p in dec.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        g_true = dec(0.5 * torch.randn(K, 8, device=dev))
        g_true = (g_true[:, 0] if g_true.dim() == 5 else g_true).contiguous()
        obs = render_depth_batched(g_true, truth["position"], truth["orientation"], truth["inv_scale"],
                                   THR, cam).contiguous()
    net_b, net_h = PointNetBackbone().to(dev).eval(), PoseHead().to(dev).eval()

    def initialise():
        """Observation -> initial (latent, position, scale, orientation) per instance: clouds, centroid
        normalisation (simple_setup.py:777-794), the network, and the squashing of its raw outputs."""
        clouds, counts = depth_to_pointclouds(obs, cam)
        n = counts.clamp(min=1)
        valid = torch.arange(clouds.shape[1], device=dev)[None] < counts[:, None]
        centroid = (clouds * valid[..., None]).sum(1) / n[:, None]
        pick = (torch.rand(K, NPTS, device=dev) * n[:, None]).long().clamp(max=clouds.shape[1] - 1)
        inp = torch.gather(clouds, 1, pick[..., None].expand(-1, -1, 3)) - centroid[:, None]
        with torch.no_grad():
            z, p, s, q = net_h(net_b(inp))
        position = centroid + 0.02 * torch.tanh(p)
        position[:, 2] -= 0.05  # the centroid of the visible surface lies in front of the object centre
        scale = 0.15 * torch.exp(0.2 * torch.tanh(s))
        return 0.1 * torch.tanh(z), position, scale, q

    init_ms = timed(initialise, 5, warm=2)
    latent, position, scale, orientation = initialise()
    opt = HypothesisOptimizer(cam, THR, obs, position, orientation, scale, latent=latent, decoder=dec,
                              optimizer="fused", inlier_threshold=0.03)  # all observed points, as the reference
    first = opt.step().clone()
    opt.capture(warmup=2)
    sizes = [K] * world

    def iteration():
        return gather_losses(opt.step(), sizes=sizes)

    for _ in range(3):
        iteration()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(ITER):
        allv = iteration()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b), init_ms], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    loop_ms, init_ms = float(ms[0]), float(ms[1])
    return {"workload": f"C5: {TOTAL} object instances ({K} per GPU), one 640x480 Redwood-shaped observation "
                        f"each; random-init PointNet + pose head (PyTorch), then {ITER} fused "
                        "render-and-compare iterations with the decoder inside (pose + scale + latent), "
                        "all_gather of the per-instance losses every iteration",
            "instances_total": TOTAL, "instances_per_gpu": K, "iterations": ITER, "scaling": "weak",
            "init_ms_per_batch": init_ms, "ms_per_iteration": loop_ms / ITER,
            "instance_iter_per_s": TOTAL * ITER / (loop_ms * 1e-3),
            "end_to_end_ms_per_batch": init_ms + loop_ms,
            "instances_per_s_end_to_end": TOTAL / ((init_ms + loop_ms) * 1e-3),
            "points_per_instance": [int(opt.point_counts.min()), int(opt.point_counts.max())],
            "mean_loss_first": float(first.mean()), "mean_loss_last": float(torch.nan_to_num(allv).mean()),
            "mean_inlier_ratio_last": float(torch.nan_to_num(opt.inlier_ratio).mean()),
            "mean_best_inlier_ratio": float(torch.nan_to_num(opt.best_inlier_ratio).mean()),
            "init": "position = cloud centroid + 0.02 tanh(net) (-0.05 m in z), scale = 0.15 exp(0.2 tanh(net)), "
                    "latent = 0.1 tanh(net), orientation = the network's unit quaternion"}


if rank == 0 and os.environ.get("SKIP_C1", "0") != "1":
    out["c1"] = config1()
if rank == 0 and os.environ.get("SKIP_C3", "0") != "1":
    out["c3"] = config3()
if world > 1:
    dist.barrier()
if os.environ.get("SKIP_C5", "0") != "1":
    out["c5"] = config5()
if rank == 0:
    tag = sys.argv[1] if len(sys.argv) > 1 else "cfg"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_configs_n{world}.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
if world > 1:
    dist.destroy_process_group()
