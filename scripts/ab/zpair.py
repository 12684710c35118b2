"""A/B of the z-pair grid layout (float2 per voxel, 4 x LDG.64 per trilinear sample) against the skewed fp32
layout (8 x LDG.32) in the REAL C2 kernels: sdfr_compare_forward and sdfr_compare_fused, 64 hypotheses x
640x480, 64^3, empty-space bounds, L2 flushed; plus the cost of producing each layout from dense grids."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import c2_case as c  # noqa: E402

lib, _lib = c.lib, c._lib
n = ctypes.c_longlong(0)
_lib.check(lib.sdfr_zpair_elems(c.R, ctypes.byref(n)), "zpair elems")
ZP = int(n.value)
zpair = torch.empty(c.B, ZP, device=c.dev)
_lib.check(lib.sdfr_zpair_grids(c.grids.data_ptr(), c.R, c.RRR, c.B, zpair.data_ptr(), ZP, c.st), "zpair")


def run(fn_name, src, stride, layout, flags=None):
    flags = (_lib.GRAD_ALL | _lib.ZERO_GRADS) if flags is None else flags
    common = (src.data_ptr(), c.R, stride, layout, c.pos.data_ptr(), c.quat.data_ptr(), c.inv_s.data_ptr(), c.B, c.W,
              c.H, 320.0, 240.0, 320.0, 320.0, c.THR, c.obs.data_ptr(), 0, c.depth.data_ptr(), c.sums[0].data_ptr(),
              c.sums[1].data_ptr())
    if fn_name == "fwd":
        _lib.check(lib.sdfr_compare_forward(*common, _lib.ZERO_GRADS, c.bounds.data_ptr(), c.st), "fwd")
    else:
        _lib.check(lib.sdfr_compare_fused(*common, c.g_sdf.data_ptr(), c.RRR, c.g_pos.data_ptr(), c.g_quat.data_ptr(),
                                          c.g_is.data_ptr(), flags, c.bounds.data_ptr(), c.st), "fused")


out = {"zpair_floats_per_grid": ZP, "skewed_floats_per_grid": c.SK}
res = {}
for name, (src, stride, layout) in (("skewed", (c.skewed, c.SK, 1)), ("zpair", (zpair, ZP, 2))):
    run("fused", src, stride, layout)
    torch.cuda.synchronize()
    res[name] = [t.clone() for t in (c.depth, c.sums, c.g_sdf, c.g_pos, c.g_quat, c.g_is)]
    out[name] = {"fwd_us": c.timed(lambda: run("fwd", src, stride, layout))["median_us"],
                 "fused_us": c.timed(lambda: run("fused", src, stride, layout))["median_us"],
                 # what the pose-only loops (fixed grids: C4 sweep, loop.pose_only) launch
                 "fused_pose_only_us": c.timed(lambda: run("fused", src, stride, layout, 0x0E | _lib.ZERO_GRADS))["median_us"]}
out["identical_depth"] = bool(torch.equal(res["skewed"][0], res["zpair"][0]))
out["identical_counts"] = bool(torch.equal(res["skewed"][1][1], res["zpair"][1][1]))
out["max_rel_grad_diff"] = max(float((a - b).abs().max() / a.abs().max().clamp(min=1e-30))
                               for a, b in zip(res["skewed"][2:], res["zpair"][2:]))
out["layout_pass_us"] = {
    "skew": c.timed(lambda: lib.sdfr_skew_grids(c.grids.data_ptr(), c.R, c.RRR, c.B, c.skewed.data_ptr(), c.SK, c.st))["median_us"],
    "zpair": c.timed(lambda: lib.sdfr_zpair_grids(c.grids.data_ptr(), c.R, c.RRR, c.B, zpair.data_ptr(), ZP, c.st))["median_us"]}
print(json.dumps(out, indent=1))
tag = sys.argv[1] if len(sys.argv) > 1 else "ab"
json.dump(out, open(os.path.join(c.ROOT, "gpurun_out", f"{tag}_zpair_ab.json"), "w"), indent=1)
