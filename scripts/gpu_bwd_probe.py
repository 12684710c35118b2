"""Probe of the unfused batched backward (C2 sizes): both / sdf-only / pose-only gradients."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200 import _lib  # noqa: E402
from sdfest_b200 import synthetic as syn  # noqa: E402

W, H, R, THR, B = 640, 480, 64, 0.005, 64
dev = torch.device("cuda:0")
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
hyp = syn.make_hypotheses(B, seed=0, device=dev)
grids = syn.hypothesis_grids(hyp["shape_param"], R, dev)
pos, quat, inv_s = hyp["position"], hyp["orientation"], hyp["inv_scale"]
depth = torch.empty(B, H, W, device=dev)
sums = torch.zeros(2, B, device=dev)
obs = torch.empty(H, W, device=dev)
RRR = R ** 3
cam = (W, H, 320.0, 240.0, 320.0, 320.0)
lib.sdfr_forward(grids.data_ptr(), R, 0, 0, pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(), 1, *cam, THR, obs.data_ptr(), None, st)
g_sdf = torch.empty_like(grids)
g_pos, g_quat, g_is = torch.empty_like(pos), torch.empty_like(quat), torch.empty_like(inv_s)
lib.sdfr_compare_forward(grids.data_ptr(), R, RRR, 0, pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(), B, *cam, THR,
                         obs.data_ptr(), 0, depth.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), _lib.ZERO_GRADS, None, st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for flags in (0x0F, 0x01, 0x0E):
    for _ in range(2):
        flush.zero_()
        lib.sdfr_compare_backward(depth.data_ptr(), obs.data_ptr(), 0, sums[1].data_ptr(), None, grids.data_ptr(), R, RRR, 0,
                                  pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(), B, *cam, g_sdf.data_ptr(), RRR,
                                  g_pos.data_ptr(), g_quat.data_ptr(), g_is.data_ptr(), flags | _lib.ZERO_GRADS, None, st)
torch.cuda.synchronize()
print("ok")
