"""C2 loop timing: B hypotheses x 640x480, Adam on pose/scale(/latent): ms per iteration, eager and
as a replayed CUDA graph, with a per-phase breakdown.  Writes gpurun_out/<tag>_loop.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200 import synthetic as syn  # noqa: E402
from sdfest_b200.differentiable_renderer import Camera, render_and_compare, render_depth_batched  # noqa: E402
from sdfest_b200.estimation import (FusedTailDecoder, HypothesisOptimizer, SDFDecoder,  # noqa: E402
                                    SurfaceDecoder, losses)

W, H, R, THR = 640, 480, 64, 0.005
B = int(os.environ.get("LOOP_B", "64"))
dev = torch.device("cuda:0")
cam = Camera(W, H, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
torch.manual_seed(0)
hyp = syn.make_hypotheses(B, seed=0, device=dev)
grids = syn.hypothesis_grids(hyp["shape_param"], R, dev)
base = syn.make_hypotheses(1, seed=0, device=dev)
obs = render_depth_batched(syn.hypothesis_grids(base["shape_param"], R, dev), base["position"],
                           base["orientation"], base["inv_scale"], THR, cam)[0].contiguous()
out = {"B": B}


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def make(decoder, optimizer="auto"):
    kw = dict(sdf=grids) if decoder is None else dict(latent=torch.zeros(B, 8, device=dev), decoder=decoder)
    return HypothesisOptimizer(cam, THR, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                               optimizer=optimizer, **kw)


torch.manual_seed(0)
_plain = SDFDecoder(R)
_fused = FusedTailDecoder(SDFDecoder(R), trunk="torch")
_fused.decoder.load_state_dict(_plain.state_dict())
ALL = os.environ.get("LOOP_ALL", "0") == "1"  # also the slow historical variants
variants = [("pose_only_torch_adam", None, "torch"), ("pose_only", None, "fused")]
if ALL:
    variants += [
        ("pose_latent", SurfaceDecoder(syn.sdf_mug(R, dev), decoder=_plain).to(dev).eval(), "torch"),
        ("pose_latent_fused_tail", SurfaceDecoder(syn.sdf_mug(R, dev), decoder=_fused).to(dev).eval(), "torch"),
        ("pose_latent_fused_iteration_torch_trunk",
         syn.residual_decoder(R, dev, syn.sdf_mug(R, dev), trunk="torch"), "torch")]
variants += [("pose_latent_fused_iteration_torch_adam", syn.residual_decoder(R, dev, syn.sdf_mug(R, dev)), "torch"),
             ("pose_latent_fused_iteration", syn.residual_decoder(R, dev, syn.sdf_mug(R, dev)), "fused")]
for name, dec, optimizer in variants:
    if dec is not None:
        for p in dec.parameters():
            p.requires_grad_(False)
    opt = make(dec, optimizer)
    eager = timeit(opt.step)
    l0 = float(opt.last_losses.mean())
    opt2 = make(dec, optimizer)
    try:
        opt2.capture()
        graph = timeit(opt2.step, n=50)
        l1 = float(opt2.last_losses.mean())
    except Exception as e:  # noqa: BLE001
        graph, l1 = None, str(e)[:200]
    out[name] = {"eager_ms": eager, "graph_ms": graph, "loss_eager": l0, "loss_graph": l1,
                 "hyp_iter_per_s_graph": (B / (graph * 1e-3)) if graph else None}

# phase breakdown (eager, decoder variant)
dec = SurfaceDecoder(syn.sdf_mug(R, dev)).to(dev).eval()
lat = torch.zeros(B, 8, device=dev, requires_grad=True)
pos = hyp["position"].clone().requires_grad_(True)
quat = hyp["orientation"].clone().requires_grad_(True)
scale = (1.0 / hyp["inv_scale"]).clone().requires_grad_(True)
pts = losses.depth_to_pointcloud(obs, cam)
pts = pts.contiguous()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


acc = {}
for it in range(8):
    e0 = ev()
    g = dec(lat)[:, 0].contiguous()
    e1 = ev()
    loss_d, _, _ = render_and_compare(g, pos, quat, (1.0 / scale).contiguous(), obs, THR, cam)
    e2 = ev()
    pc = losses.point_loss(pts, pos, quat, scale, g)
    e3 = ev()
    (torch.nan_to_num(loss_d) + 3.0 * pc).sum().backward()
    e4 = ev()
    torch.cuda.synchronize()
    if it >= 3:
        for k, a, b in (("decode", e0, e1), ("render_compare_fwd", e1, e2), ("pc_loss_fwd", e2, e3),
                        ("backward_all", e3, e4)):
            acc[k] = acc.get(k, 0.0) + a.elapsed_time(b) / 5
    for t in (lat, pos, quat, scale):
        t.grad = None
out["phases_ms_eager"] = acc
tag = sys.argv[1] if len(sys.argv) > 1 else "loop"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_loop.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
