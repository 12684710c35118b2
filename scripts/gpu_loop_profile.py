"""Per-kernel CUDA time of ONE loop iteration (HypothesisOptimizer.step, eager) for the decoder
variants: which kernels are left once the renderer and the decoder tail are fused.
Writes gpurun_out/<tag>_loop_ops_<variant>.txt."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200 import synthetic as syn  # noqa: E402
from sdfest_b200.differentiable_renderer import Camera, render_depth_batched  # noqa: E402
from sdfest_b200.estimation import (FusedTailDecoder, HypothesisOptimizer, SDFDecoder,  # noqa: E402
                                    SurfaceDecoder)

W, H, R, THR = 640, 480, 64, 0.005
B = int(os.environ.get("LOOP_B", "64"))
variants = (os.environ.get("LOOP_VARIANTS") or "fused_tail").split(",")
dev = torch.device("cuda:0")
cam = Camera(W, H, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
hyp = syn.make_hypotheses(B, seed=0, device=dev)
base = syn.make_hypotheses(1, seed=0, device=dev)
obs = render_depth_batched(syn.hypothesis_grids(base["shape_param"], R, dev), base["position"],
                           base["orientation"], base["inv_scale"], THR, cam)[0].contiguous()
tag = sys.argv[1] if len(sys.argv) > 1 else "prof"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for v in variants:
    torch.manual_seed(0)
    d = SDFDecoder(R)
    if v == "fused_tail":
        d = FusedTailDecoder(d)
    dec = SurfaceDecoder(syn.sdf_mug(R, dev), decoder=d).to(dev).eval()
    if v == "fused_iteration":
        dec = syn.residual_decoder(R, dev, syn.sdf_mug(R, dev))
    for p in dec.parameters():
        p.requires_grad_(False)
    opt = HypothesisOptimizer(cam, THR, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                              latent=torch.zeros(B, 8, device=dev), decoder=dec)
    for _ in range(5):
        opt.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            opt.step()
        torch.cuda.synchronize()
    txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=110)
    open(os.path.join(ROOT, "gpurun_out", f"{tag}_loop_ops_{v}.txt"), "w").write(txt)
    print("=====", v)
    print(txt)
