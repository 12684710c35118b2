"""Debug: C3 pose gradients -- composite backward vs single-object backward vs oracle, per object."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from sdfest_b200 import _lib
from sdfest_b200.differentiable_renderer import Camera, render_depth_composite
from test_bench_parity import c3_scene
dev = torch.device("cuda:0")
lib = _lib.lib()
W, H, THR, R = 1280, 720, 0.005, 128
cam_d = dict(cx=640.0, cy=360.0, fx=640.0, fy=640.0)
cam = Camera(W, H, 640.0, 640.0, 640.0, 360.0, pixel_center=0.5)
grids, pos, quat, inv_s = c3_scene(dev)
a = [t.clone().requires_grad_(True) for t in (grids, pos, quat, inv_s)]
depth, winner = render_depth_composite(*a, THR, cam)
g = np.random.default_rng(4).standard_normal((H, W)).astype(np.float32)
tg = torch.as_tensor(g, device=dev)
depth.backward(tg)
st = torch.cuda.current_stream().cuda_stream
for k in (2, 7, 0, 5):
    dk = torch.where(winner == k, depth.detach(), torch.zeros_like(depth)).contiguous()
    gp, gq, gi = torch.empty(3, device=dev), torch.empty(4, device=dev), torch.empty(1, device=dev)
    _lib.check(lib.sdfr_backward(tg.data_ptr(), dk.data_ptr(), grids[k].data_ptr(), R, 0, 0, pos[k].data_ptr(),
                                 quat[k].data_ptr(), inv_s[k:k + 1].data_ptr(), 1, W, H, 640.0, 360.0, 640.0, 640.0,
                                 None, 0, gp.data_ptr(), gq.data_ptr(), gi.data_ptr(), 0x0E | _lib.ZERO_GRADS, None, st), "bwd")
    torch.cuda.synchronize()
    bw = oracle.render_backward(g, dk.cpu().numpy(), grids[k].cpu().numpy(), pos[k].cpu().numpy(), quat[k].cpu().numpy(),
                                inv_s[k:k + 1].cpu().numpy(), W, H, want_sdf=False, nthreads=16, **cam_d)
    bw64 = oracle.render_backward(g.astype(np.float64), dk.cpu().numpy().astype(np.float64), grids[k].cpu().numpy(), pos[k].cpu().numpy(), quat[k].cpu().numpy(),
                                inv_s[k:k + 1].cpu().numpy(), W, H, want_sdf=False, nthreads=16, dtype=np.float64, **cam_d)
    print(f"obj {k}: n={(dk>0).sum().item()}")
    print("  composite ", a[2].grad[k].cpu().numpy(), a[3].grad[k].item())
    print("  single    ", gq.cpu().numpy(), gi.item())
    print("  oracle32  ", bw["g_orientation"], bw["g_inv_scale"])
    print("  oracle64  ", bw64["g_orientation"], bw64["g_inv_scale"])
