"""Shared setup of the C2 kernel A/B scripts: 64 hypotheses x 640x480, 64^3 skewed grids, bounds, timers."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sdfest_b200 import _lib  # noqa: E402
from sdfest_b200 import synthetic as syn  # noqa: E402

W, H, R, THR, B = 640, 480, 64, 0.005, 64
RRR = R ** 3
dev = torch.device("cuda:0")
lib = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
hyp = syn.make_hypotheses(B, seed=0, device=dev)
grids = syn.hypothesis_grids(hyp["shape_param"], R, dev)
pos, quat, inv_s = hyp["position"], hyp["orientation"], hyp["inv_scale"]
depth = torch.empty(B, H, W, device=dev)
sums = torch.zeros(2, B, device=dev)
obs = torch.empty(H, W, device=dev)
lib.sdfr_forward(grids.data_ptr(), R, 0, 0, pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(), 1, W, H,
                 320.0, 240.0, 320.0, 320.0, THR, obs.data_ptr(), None, st)
g_sdf = torch.empty_like(grids)
g_pos, g_quat, g_is = torch.empty_like(pos), torch.empty_like(quat), torch.empty_like(inv_s)
n_sk = ctypes.c_longlong(0)
lib.sdfr_skewed_pitches(R, None, None, ctypes.byref(n_sk))
SK = int(n_sk.value)
skewed = torch.empty(B, SK, device=dev)
lib.sdfr_skew_grids(grids.data_ptr(), R, RRR, B, skewed.data_ptr(), SK, st)
bounds = torch.empty(B, 8, dtype=torch.int32, device=dev)
lib.sdfr_grid_bounds(skewed.data_ptr(), R, SK, 1, pos.data_ptr(), inv_s.data_ptr(), B, THR, bounds.data_ptr(), st)


def timed(fn, n=30, warm=5, do_flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return {"median_us": round(ts[len(ts) // 2], 2), "min_us": round(ts[0], 2)}


def fwd():
    lib.sdfr_compare_forward(skewed.data_ptr(), R, SK, 1, pos.data_ptr(), quat.data_ptr(),
                             inv_s.data_ptr(), B, W, H, 320.0, 240.0, 320.0, 320.0, THR,
                             obs.data_ptr(), 0, depth.data_ptr(), sums[0].data_ptr(),
                             sums[1].data_ptr(), _lib.ZERO_GRADS, bounds.data_ptr(), st)


def fused(flags=_lib.GRAD_ALL | _lib.ZERO_GRADS):
    lib.sdfr_compare_fused(skewed.data_ptr(), R, SK, 1, pos.data_ptr(), quat.data_ptr(),
                           inv_s.data_ptr(), B, W, H, 320.0, 240.0, 320.0, 320.0, THR,
                           obs.data_ptr(), 0, depth.data_ptr(), sums[0].data_ptr(),
                           sums[1].data_ptr(), g_sdf.data_ptr(), RRR, g_pos.data_ptr(),
                           g_quat.data_ptr(), g_is.data_ptr(), flags, bounds.data_ptr(), st)


def bwd(flags=_lib.GRAD_ALL | _lib.ZERO_GRADS):
    lib.sdfr_compare_backward(depth.data_ptr(), obs.data_ptr(), 0, sums[1].data_ptr(), None,
                              skewed.data_ptr(), R, SK, 1, pos.data_ptr(), quat.data_ptr(),
                              inv_s.data_ptr(), B, W, H, 320.0, 240.0, 320.0, 320.0,
                              g_sdf.data_ptr(), RRR, g_pos.data_ptr(), g_quat.data_ptr(),
                              g_is.data_ptr(), flags, bounds.data_ptr(), st)
