"""sdfr_decoder_tail_forward + sdfr_grid_bounds against sdfr_decoder_tail_forward_bounds (64 x 4x30^3 -> 64^3 skewed)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sdfest_b200 import _lib, synthetic as syn
dev = torch.device("cuda:0"); lib = _lib.lib()
B, C, S, R, THR = 64, 4, 30, 64, 0.005
x = torch.randn(B, C, S, S, S, device=dev) * 0.05; w = torch.randn(C, device=dev); b = torch.zeros(1, device=dev)
base = syn.sdf_mug(R, dev).contiguous()
hyp = syn.make_hypotheses(B, seed=0, device=dev)
n = ctypes.c_longlong(0); lib.sdfr_skewed_pitches(R, None, None, ctypes.byref(n)); SK = int(n.value)
g = torch.empty(B, SK, device=dev); bounds = torch.empty(B, 8, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def tail(): lib.sdfr_decoder_tail_forward(x.data_ptr(), C, S, w.data_ptr(), b.data_ptr(), base.data_ptr(), B, R, g.data_ptr(), SK, 1, st)
def scan(): lib.sdfr_grid_bounds(g.data_ptr(), R, SK, 1, hyp["position"].data_ptr(), hyp["inv_scale"].data_ptr(), B, THR, bounds.data_ptr(), st)
def both(): tail(); scan()
def fusedtb(): lib.sdfr_decoder_tail_forward_bounds(x.data_ptr(), C, S, w.data_ptr(), b.data_ptr(), base.data_ptr(), B, R, g.data_ptr(), SK, 1, hyp["position"].data_ptr(), hyp["inv_scale"].data_ptr(), THR, bounds.data_ptr(), st)
def timed(fn, do_flush):
    ts = []
    for i in range(35):
        if do_flush: flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(a.elapsed_time(e) * 1e3)
    ts.sort(); return round(ts[len(ts) // 2], 1)
out = {k: {"cold_us": timed(f, True), "warm_us": timed(f, False)} for k, f in (("tail", tail), ("scan", scan), ("tail_then_scan", both), ("tail_with_bounds", fusedtb))}
print(json.dumps(out))
