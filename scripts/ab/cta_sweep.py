"""C2 kernels against the number of CTAs per hypothesis (SDFR_TARGET_CTAS = G * 64): wave quantisation.
usage: python scripts/ab/cta_sweep.py <tag> [G ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import c2_case as c  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "cta"
Gs = [int(x) for x in sys.argv[2:]] or [11, 12, 17, 23, 24, 29, 34, 35, 37, 40, 46, 47, 58, 69]
out = {}
c.fwd()
for G in Gs:
    os.environ["SDFR_TARGET_CTAS"] = str(G * c.B)
    c.fwd()
    out[G] = {"ctas": G * c.B, "fwd": c.timed(c.fwd)["median_us"], "fused": c.timed(c.fused)["median_us"],
              "bwd": c.timed(c.bwd)["median_us"]}
    print(G, out[G], flush=True)
os.makedirs(os.path.join(c.ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(c.ROOT, "gpurun_out", f"{tag}_cta_sweep.json"), "w"), indent=1)
