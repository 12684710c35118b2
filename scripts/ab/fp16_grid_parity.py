"""North-star option n1 (fp16 grid staged in shared memory for <= 48^3): what rounding the grid to fp16 does
to the depth image, measured with the product's own fp32 kernels on the fp16-rounded values -- an upper
bound on the quality of ANY fp16-grid kernel.  The parity bar is 1e-5 relative (BASELINE.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sdfest_b200 import synthetic as syn  # noqa: E402
from sdfest_b200.differentiable_renderer import Camera, render_depth_batched  # noqa: E402

dev = torch.device("cuda:0")
W, H, THR, B = 640, 480, 0.005, 16
cam = Camera(W, H, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
out = {}
for R in (32, 48):
    hyp = syn.make_hypotheses(B, seed=0, device=dev)
    grids = syn.hypothesis_grids(hyp["shape_param"], R, dev)
    d32 = render_depth_batched(grids, hyp["position"], hyp["orientation"], hyp["inv_scale"], THR, cam)
    for name, q in (("fp16", grids.half().float()), ("bf16", grids.bfloat16().float())):
        d16 = render_depth_batched(q.contiguous(), hyp["position"], hyp["orientation"], hyp["inv_scale"], THR, cam)
        both = (d32 > 0) & (d16 > 0)
        rel = ((d32 - d16).abs() / d32.clamp(min=1e-9))[both]
        out[f"R{R}_{name}"] = {
            "hit_pixels": int((d32 > 0).sum()), "mask_flips": int(((d32 > 0) != (d16 > 0)).sum()),
            "frac_hits_beyond_1e-5": float((rel > 1e-5).float().mean()),
            "median_rel": float(rel.median()), "p99_rel": float(rel.quantile(0.99)), "max_rel": float(rel.max())}
print(json.dumps(out, indent=1))
tag = sys.argv[1] if len(sys.argv) > 1 else "ab"
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_fp16_grid_parity.json"), "w"), indent=1)
