import cProfile, pstats, sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sdfest_b200 import synthetic as syn
from sdfest_b200.differentiable_renderer import Camera, render_depth_gpu
dev = torch.device("cuda:0")
W, H, R, THR = 640, 480, 64, 0.005
cam = Camera(W, H, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
hyp = syn.make_hypotheses(1, seed=0, device=dev)
grid = syn.hypothesis_grids(hyp["shape_param"], R, dev)[0].contiguous()
p, q, s = hyp["position"][0].clone(), hyp["orientation"][0].clone(), hyp["inv_scale"].clone()
g = torch.randn(H, W, device=dev)
def once():
    a = [grid.detach().requires_grad_(True), p.detach().requires_grad_(True), q.detach().requires_grad_(True), s.detach().requires_grad_(True)]
    d = render_depth_gpu(*a, threshold=THR, camera=cam)
    d.backward(g)
for _ in range(200): once()
torch.cuda.synchronize()
N = 3000
t0 = time.perf_counter()
for _ in range(N): once()
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
print("cpu-side us per fwd+bwd (launch only):", t_cpu / N * 1e6, "total incl. drain:", (time.perf_counter() - t0) / N * 1e6)
def leaves_only():
    a = [grid.detach().requires_grad_(True), p.detach().requires_grad_(True), q.detach().requires_grad_(True), s.detach().requires_grad_(True)]
t0 = time.perf_counter()
for _ in range(N): leaves_only()
print("leaf creation us:", (time.perf_counter() - t0) / N * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(N): once()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
