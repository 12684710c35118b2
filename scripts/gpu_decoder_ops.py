"""Per-launch CUDA-event time of every decoder kernel at the C2 shapes (B hypotheses, mug decoder):
conv forward / dgrad of the three stages, the two resizes and the tail, forward and adjoint.
Writes gpurun_out/<tag>_decoder_ops.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200 import _lib  # noqa: E402

B = int(os.environ.get("LOOP_B", "64"))
dev = torch.device("cuda:0")
lib = _lib.lib()
torch.manual_seed(0)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


N = int(os.environ.get("OPS_N", "20"))  # OPS_N=1: one launch per kernel (ncu capture)


def timeit(fn, n=N):
    for _ in range(3 if n > 1 else 0):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = {"B": B}
# (in_size U, Ci, Co): conv U -> U-2
for U, Ci, Co in ((8, 16, 16), (16, 16, 8), (32, 8, 4)):
    O = U - 2
    x = torch.randn(B, Ci, U, U, U, device=dev)
    w = torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.1
    bias = torch.randn(Co, device=dev)
    y = torch.empty(B, Co, O, O, O, device=dev)
    gy = torch.randn_like(y)
    gx = torch.empty_like(x)
    flops = 2 * 27 * Ci * Co * O ** 3 * B
    t = timeit(lambda: _lib.check(lib.sdfr_conv3d_forward(x.data_ptr(), B, Ci, U, w.data_ptr(), bias.data_ptr(),
                                                          Co, 3, 1, y.data_ptr(), None), "conv"))
    out[f"conv_fwd_{Ci}to{Co}_{U}"] = {"us": t, "tflops": flops / t * 1e-6}
    t = timeit(lambda: _lib.check(lib.sdfr_conv3d_backward_data(gy.data_ptr(), y.data_ptr(), B, Ci, U,
                                                                w.data_ptr(), Co, 3, gx.data_ptr(), None), "dgrad"))
    out[f"conv_dgrad_{Ci}to{Co}_{U}"] = {"us": t, "tflops": flops / t * 1e-6}
for S, U, C in ((6, 16, 16), (14, 32, 8)):
    x = torch.randn(B * C, S, S, S, device=dev)
    y = torch.empty(B * C, U, U, U, device=dev)
    gx = torch.empty_like(x)
    byts = 4 * (x.numel() + y.numel())
    t = timeit(lambda: _lib.check(lib.sdfr_upsample3d_forward(x.data_ptr(), B * C, S, U, y.data_ptr(), None), "up"))
    out[f"up_fwd_{S}to{U}_c{C}"] = {"us": t, "GBps": byts / t * 1e-3}
    t = timeit(lambda: _lib.check(lib.sdfr_upsample3d_backward(y.data_ptr(), B * C, S, U, gx.data_ptr(), None), "upb"))
    out[f"up_bwd_{S}to{U}_c{C}"] = {"us": t, "GBps": byts / t * 1e-3}
S, R, C = 30, 64, 4
x = torch.randn(B, C, S, S, S, device=dev)
w = torch.randn(C, device=dev)
bias = torch.randn(1, device=dev)
g = torch.randn(B, R ** 3, device=dev)
g2 = torch.randn(B, R ** 3, device=dev)
n = torch.full((B,), 100.0, device=dev)
up = torch.ones(B, device=dev)
gx = torch.empty_like(x)
from sdfest_b200.differentiable_renderer.sdf_renderer import _skewed_elems  # noqa: E402

SK = _skewed_elems(R)
sk = torch.empty(B, SK, device=dev)
byts = 4 * (x.numel() + B * R ** 3)
t = timeit(lambda: _lib.check(lib.sdfr_decoder_tail_forward(x.data_ptr(), C, S, w.data_ptr(), bias.data_ptr(), None,
                                                            B, R, sk.data_ptr(), SK, 1, None), "tail"))
out["tail_fwd_skewed"] = {"us": t, "GBps": byts / t * 1e-3}
t = timeit(lambda: _lib.check(lib.sdfr_decoder_tail_backward(g.data_ptr(), R ** 3, n.data_ptr(), up.data_ptr(),
                                                             g2.data_ptr(), R ** 3, w.data_ptr(), C, S, B, R,
                                                             gx.data_ptr(), None), "tailb"))
out["tail_bwd_two_grids"] = {"us": t, "GBps": (byts + 4 * B * R ** 3) / t * 1e-3}
t = timeit(lambda: _lib.check(lib.sdfr_decoder_tail_backward(g.data_ptr(), R ** 3, n.data_ptr(), up.data_ptr(),
                                                             None, 0, w.data_ptr(), C, S, B, R,
                                                             gx.data_ptr(), None), "tailb"))
out["tail_bwd_one_grid"] = {"us": t, "GBps": byts / t * 1e-3}
tag = sys.argv[1] if len(sys.argv) > 1 else "ops"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_decoder_ops.json"), "w"), indent=1)
for k, v in out.items():
    print(k, v)
