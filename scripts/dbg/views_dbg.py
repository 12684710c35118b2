import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_views as tv
from sdfest_b200.estimation import HypothesisOptimizer
dev = torch.device("cuda:0")
B = 5
cam, thr, obs, hyp, cam_p, cam_q, kw = tv._view_scene(dev, B, 32, False, V=3)
for pcw in (3.0, 0.0):
    def make(optimizer):
        return HypothesisOptimizer(cam, thr, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                                   camera_positions=cam_p, camera_orientations=cam_q, inlier_threshold=0.03,
                                   max_points=1500, optimizer=optimizer, pc_weight=pcw, **kw)
    a, b = make("torch"), make("fused")
    for it in range(3):
        pa = [t.detach().clone() for t in (a.position, a.orientation, a.scale)]
        pb = [t.detach().clone() for t in (b.position, b.orientation, b.scale)]
        la, lb = a.step().clone(), b.step().clone()
        print("pcw", pcw, "it", it, "\n torch", la.tolist(), "\n fused", lb.tolist())
        for n, x, y, x0, y0 in zip("pqs", (a.position, a.orientation, a.scale), (b.position, b.orientation, b.scale), pa, pb):
            print("  d", n, (x.detach() - x0)[0].tolist(), (y.detach() - y0)[0].tolist())
