"""ms per iteration of the fused view loop (V views of B hypotheses, decoder inside, CUDA-graph replay).
usage: python scripts/gpu_views_timing.py [tag] -> gpurun_out/<tag>_views.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_views as tv  # noqa: E402
from sdfest_b200.estimation import HypothesisOptimizer  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for B, V in ((64, 3), (64, 2), (1, 3)):
    cam, thr, obs, hyp, cam_p, cam_q, kw = tv._view_scene(dev, B, 64, True, V=V)
    opt = HypothesisOptimizer(cam, thr, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                              camera_positions=cam_p, camera_orientations=cam_q, inlier_threshold=0.03,
                              max_points=20000, optimizer="fused", **kw)
    opt.capture(warmup=3)
    for _ in range(5):
        opt.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        opt.step()
    b.record()
    torch.cuda.synchronize()
    out[f"B{B}_V{V}"] = {"ms_per_iteration": a.elapsed_time(b) / 50, "camera": [cam.width, cam.height]}
tag = sys.argv[1] if len(sys.argv) > 1 else "views"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_views.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
