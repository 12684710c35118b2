"""C1 (one 640x480 / 64^3 frame, forward + backward): a few C-ABI iterations for an ncu launch list.
usage: ncu --metrics gpu__time_duration.sum ... python scripts/gpu_c1_probe.py [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200 import _lib  # noqa: E402
from sdfest_b200 import synthetic as syn  # noqa: E402

W, H, R, THR = 640, 480, 64, 0.005
dev = torch.device("cuda:0")
lib = _lib.lib()
hyp = syn.make_hypotheses(1, seed=0, device=dev)
grid = syn.hypothesis_grids(hyp["shape_param"], R, dev)[0].contiguous()
p, q, s = hyp["position"][0].clone(), hyp["orientation"][0].clone(), hyp["inv_scale"].clone()
g = torch.randn(H, W, device=dev)
depth = torch.empty(1, H, W, device=dev)
gs, gp, gq, gi = torch.empty_like(grid), torch.empty(3, device=dev), torch.empty(4, device=dev), torch.empty(1, device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    flush.zero_()
    lib.sdfr_forward(grid.data_ptr(), R, 0, 0, p.data_ptr(), q.data_ptr(), s.data_ptr(), 1, W, H,
                     320.0, 240.0, 320.0, 320.0, THR, depth.data_ptr(), None, st)
    lib.sdfr_backward(g.data_ptr(), depth.data_ptr(), grid.data_ptr(), R, 0, 0, p.data_ptr(),
                      q.data_ptr(), s.data_ptr(), 1, W, H, 320.0, 240.0, 320.0, 320.0,
                      gs.data_ptr(), 0, gp.data_ptr(), gq.data_ptr(), gi.data_ptr(),
                      _lib.GRAD_ALL | _lib.ZERO_GRADS, None, st)
torch.cuda.synchronize()
print("hit pixels", int((depth > 0).sum()))
