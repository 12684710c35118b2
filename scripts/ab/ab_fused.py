"""Same-box A/B of two builds of libsdfrender.so on the C2 kernels (SDFR_LIB_PATH selects the build).
usage: python scripts/ab/ab_fused.py  -> prints median us of fused / forward for the loaded build."""
import os, sys, ctypes, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sdfest_b200 import _lib
from sdfest_b200 import synthetic as syn
from sdfest_b200.differentiable_renderer import Camera, render_depth_batched
W, H, R, THR, B = 640, 480, 64, 0.005, 64
dev = torch.device("cuda:0")
cam = Camera(W, H, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
# raw load: the other build may have an older ABI (fewer symbols) -- bind what it exports
lib = ctypes.CDLL(_lib.LIB_PATH)
for _name, (_res, _args) in _lib.SIGNATURES.items():
    if hasattr(lib, _name):
        getattr(lib, _name).restype, getattr(lib, _name).argtypes = _res, _args
_lib._lib = lib  # render_depth_batched below goes through the same handle
hyp = syn.make_hypotheses(B, seed=0, device=dev)
grids = syn.hypothesis_grids(hyp["shape_param"], R, dev)
pos, quat, inv_s = hyp["position"], hyp["orientation"], hyp["inv_scale"]
base = syn.make_hypotheses(1, seed=0, device=dev)
obs = render_depth_batched(syn.hypothesis_grids(base["shape_param"], R, dev), base["position"], base["orientation"], base["inv_scale"], THR, cam)[0].contiguous()
n_sk = ctypes.c_longlong(0); lib.sdfr_skewed_pitches(R, None, None, ctypes.byref(n_sk)); SK = int(n_sk.value)
skewed = torch.empty(B, SK, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.sdfr_skew_grids(grids.data_ptr(), R, R**3, B, skewed.data_ptr(), SK, st)
depth = torch.empty(B, H, W, device=dev); sums = torch.zeros(2, B, device=dev)
g_sdf = torch.empty_like(grids); g_pos, g_quat, g_is = torch.empty_like(pos), torch.empty_like(quat), torch.empty_like(inv_s)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def fused():
    assert 0 == (lib.sdfr_compare_fused(skewed.data_ptr(), R, SK, 1, pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(), B, W, H, 320.0, 240.0, 320.0, 320.0, THR, obs.data_ptr(), 0, depth.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), g_sdf.data_ptr(), R**3, g_pos.data_ptr(), g_quat.data_ptr(), g_is.data_ptr(), _lib.GRAD_ALL | _lib.ZERO_GRADS, None, st))
def fwd():
    assert 0 == (lib.sdfr_forward(skewed.data_ptr(), R, SK, 1, pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(), B, W, H, 320.0, 240.0, 320.0, 320.0, THR, depth.data_ptr(), None, st))
def timed(fn, n=40):
    ts = []
    for i in range(n + 5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
print(json.dumps({"lib": os.environ.get("SDFR_LIB_PATH", "default"), "fused_us": timed(fused), "forward_us": timed(fwd), "checksum": float(depth.sum())}))
