"""Median us of the C2 kernels for the build selected by SDFR_LIB_PATH (same box A/B of build variants)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import c2_case as c  # noqa: E402

c.fwd()
out = {"lib": os.path.basename(os.environ.get("SDFR_LIB_PATH", "default")),
       "fwd": c.timed(c.fwd)["median_us"], "fused": c.timed(c.fused)["median_us"],
       "fused_pose_only": c.timed(lambda: c.fused(0x0E | c._lib.ZERO_GRADS))["median_us"],
       "fused_sdf_only": c.timed(lambda: c.fused(0x01 | c._lib.ZERO_GRADS))["median_us"],
       "bwd": c.timed(c.bwd)["median_us"], "bwd_sdf_only": c.timed(lambda: c.bwd(0x01 | c._lib.ZERO_GRADS))["median_us"]}
print(json.dumps(out))
