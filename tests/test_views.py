"""Host mathematics of the multi-view loop (sdfest_b200/estimation/views.py) on CPU: against golden
vectors of the reference's own quaternion_utils / view-loop lines, and the analytic pull-back of the
pose gradients against autograd."""
import os

import numpy as np
import torch

from sdfest_b200.estimation import views
from util import GOLDEN_DIR


def _golden():
    z = np.load(os.path.join(GOLDEN_DIR, "views.npz"))
    return {k: torch.tensor(z[k]) for k in z.files}


def test_quaternion_helpers_match_reference_golden():
    z = _golden()
    torch.testing.assert_close(views.quaternion_multiply(z["q1"], z["q2"]), z["multiply"], rtol=0, atol=1e-14)
    torch.testing.assert_close(views.quaternion_apply(z["q1"], z["points"]), z["apply"], rtol=0, atol=1e-14)
    assert torch.equal(views.quaternion_invert(z["q1"]), z["invert"])
    # broadcasting as the reference's ("normal broadcasting rules apply")
    out = views.quaternion_multiply(z["q1"][:, None], z["q2"][None])
    assert out.shape == (7, 7, 4)
    torch.testing.assert_close(out[2, 5], views.quaternion_multiply(z["q1"][2], z["q2"][5]))


def test_non_unit_quaternions_and_point_constraint_loss_match_reference_golden():
    """The reference never normalises inside quaternion_apply; the point-constraint loss is evaluated on
    the un-normalised orientation and its gradient depends on that (losses.py:138-153)."""
    z = _golden()
    torch.testing.assert_close(views.quaternion_apply(z["q_raw"], z["points"][:6]), z["apply_raw"], rtol=0, atol=1e-13)
    q = z["q_raw"].clone().requires_grad_(True)
    loss = views.point_constraint_loss(q, z["source"], z["target"])
    torch.testing.assert_close(loss.detach(), z["constraint_loss"], rtol=0, atol=1e-13)
    loss.sum().backward()
    torch.testing.assert_close(q.grad, z["constraint_grad"], rtol=0, atol=1e-12)


def test_camera_frames_match_the_reference_view_loop():
    z = _golden()
    pos_c, ori_c = views.to_camera_frames(z["position"], z["orientation"], z["camera_positions"],
                                          z["camera_orientations"])
    torch.testing.assert_close(pos_c, z["position_c"], rtol=0, atol=1e-14)
    torch.testing.assert_close(ori_c, z["orientation_c"], rtol=0, atol=1e-14)
    # an identity camera at the origin leaves the pose alone
    eye_p, eye_q = torch.zeros(1, 3, dtype=torch.float64), torch.tensor([[0.0, 0, 0, 1]], dtype=torch.float64)
    p1, q1 = views.to_camera_frames(z["position"], z["orientation"], eye_p, eye_q)
    assert torch.equal(p1[0], z["position"]) and torch.equal(q1[0], z["orientation"])


def test_pull_back_is_the_adjoint_of_the_view_maps():
    z = _golden()
    pos = z["position"].clone().requires_grad_(True)
    ori = z["orientation"].clone().requires_grad_(True)
    pos_c, ori_c = views.to_camera_frames(pos, ori, z["camera_positions"], z["camera_orientations"])
    g = torch.Generator().manual_seed(1)
    g_pc = torch.randn(pos_c.shape, generator=g, dtype=torch.float64)
    g_qc = torch.randn(ori_c.shape, generator=g, dtype=torch.float64)
    ((pos_c * g_pc).sum() + (ori_c * g_qc).sum()).backward()
    g_p, g_q = views.pull_back(g_pc, g_qc, z["camera_orientations"])
    torch.testing.assert_close(g_p, pos.grad, rtol=0, atol=1e-13)
    torch.testing.assert_close(g_q, ori.grad, rtol=0, atol=1e-13)


def test_multiview_step_plumbing_on_cpu(monkeypatch):
    """The V-view iteration of HypothesisOptimizer with the CUDA renderer replaced by a differentiable
    stand-in (the point loss is the real torch statement): every view sees the pose in ITS camera frame
    and its own observation and points, the per-view losses are summed, autograd reaches the world-frame
    parameters through the rigid maps, and one identity view reproduces the single-view optimiser."""
    import pytest

    from sdfest_b200.differentiable_renderer import Camera
    from sdfest_b200.estimation import HypothesisOptimizer, hypotheses

    seen = []

    def fake_render_and_compare(sdf, position, orientation, inv_scale, depth_obs, threshold, camera):
        seen.append((position.detach().clone(), orientation.detach().clone(), depth_obs))
        B = position.shape[0]
        target = depth_obs[depth_obs > 0].mean()
        loss = ((position[:, 2] + target) ** 2 + 0.1 * (orientation[:, 3] - 1) ** 2 + 0.01 * inv_scale)
        depth = depth_obs[None].expand(B, -1, -1) * 1.01
        return loss, depth, torch.ones(B)

    monkeypatch.setattr(hypotheses, "render_and_compare", fake_render_and_compare)
    from oracle.pc_loss import point_loss
    from sdfest_b200.estimation import losses

    monkeypatch.setattr(losses, "point_loss", point_loss)  # the product's point loss is CUDA-only
    W, H, B = 16, 12, 3
    cam = Camera(W, H, 14.0, 14.0, 8.0, 6.0, pixel_center=0.5)
    obs = torch.zeros(2, H, W)
    obs[0, 3:8, 4:10] = 0.9
    obs[1, 2:6, 5:9] = 1.1
    g = torch.Generator().manual_seed(2)
    pos = torch.tensor([[0.0, 0.0, -1.0]]) + 0.05 * torch.randn(B, 3, generator=g)
    quat = torch.nn.functional.normalize(torch.tensor([[0.0, 0, 0, 1]]) + 0.1 * torch.randn(B, 4, generator=g), dim=1)
    scale = torch.full((B,), 0.3)
    sdf = torch.rand(1, 8, 8, 8, generator=g) - 0.3
    cam_p = torch.tensor([[0.0, 0, 0], [0.2, 0.0, -0.1]])
    cam_q = torch.nn.functional.normalize(torch.tensor([[0.0, 0, 0, 1], [0.0, 0.3, 0.0, 1.0]]), dim=1)

    opt = HypothesisOptimizer(cam, 0.005, obs, pos, quat, scale, sdf=sdf, camera_positions=cam_p,
                              camera_orientations=cam_q, inlier_threshold=0.03)
    assert opt.optimizer_impl == "torch" and len(opt._view_points) == 2
    assert opt._view_points[0].shape == (30, 3) and opt._view_points[1].shape == (16, 3)
    before = [t.detach().clone() for t in (opt.position, opt.orientation, opt.scale)]
    loss = opt.step()
    assert loss.shape == (B,) and bool(torch.isfinite(loss).all())
    # the stand-in saw view 0 in the world frame and view 1 in the second camera's frame
    want_p, want_q = views.to_camera_frames(before[0], before[1], cam_p, cam_q)
    assert len(seen) == 2
    for v in range(2):
        torch.testing.assert_close(seen[v][0], want_p[v])
        torch.testing.assert_close(seen[v][1], want_q[v])
        assert torch.equal(seen[v][2], obs[v])
    for a, b in zip(before, (opt.position, opt.orientation, opt.scale)):
        assert float((a - b.detach()).abs().max()) > 0  # every group received a gradient
    torch.testing.assert_close(torch.linalg.norm(opt.orientation.detach(), dim=1), torch.ones(B))
    # inlier ratio of the LAST view (1 % error everywhere it is observed: all inliers)
    assert opt.inlier_ratio.tolist() == [1.0] * B

    # one identity view = the single-view optimiser
    seen.clear()
    a = HypothesisOptimizer(cam, 0.005, obs[:1], pos, quat, scale, sdf=sdf, camera_positions=cam_p[:1],
                            camera_orientations=cam_q[:1])
    b = HypothesisOptimizer(cam, 0.005, obs[0], pos, quat, scale, sdf=sdf, optimizer="torch")
    for _ in range(3):
        la, lb = a.step(), b.step()
        torch.testing.assert_close(la, lb)
    torch.testing.assert_close(a.position.detach(), b.position.detach())
    torch.testing.assert_close(a.orientation.detach(), b.orientation.detach())

    # point constraint (simple_setup.py:164-175): one more loss term on the un-normalised orientation
    src, tgt = torch.tensor([0.0, 1.0, 0.0]), torch.tensor([0.0, 0.0, 1.0])
    c = HypothesisOptimizer(cam, 0.005, obs[0], pos, quat, scale, sdf=sdf, point_constraint=(src, tgt, 2.0))
    d = HypothesisOptimizer(cam, 0.005, obs[0], pos, quat, scale, sdf=sdf, optimizer="torch")
    assert c.optimizer_impl == "torch"
    lc, ld = c.step(), d.step()
    torch.testing.assert_close(lc - ld, 2.0 * views.point_constraint_loss(quat, src, tgt))
    assert float((c.orientation.detach() - d.orientation.detach()).abs().max()) > 1e-4
    with pytest.raises(ValueError):
        HypothesisOptimizer(cam, 0.005, obs[0], pos, quat, scale, sdf=sdf, point_constraint=(src, tgt, 2.0),
                            optimizer="fused")

    for bad in (dict(camera_positions=cam_p), dict(camera_positions=cam_p, camera_orientations=cam_q[:1]),
                dict(camera_positions=cam_p, camera_orientations=cam_q, instance=torch.zeros(B, dtype=torch.long))):
        with pytest.raises(ValueError):
            HypothesisOptimizer(cam, 0.005, obs, pos, quat, scale, sdf=sdf, **bad)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_multiview_optimizer_on_the_gpu(cuda_device):
    """Two views of one object through the real renderer: the unperturbed hypothesis explains both
    observations, one identity view reproduces the single-view optimiser, the summed loss goes down."""
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.differentiable_renderer import Camera, render_depth_batched
    from sdfest_b200.estimation import HypothesisOptimizer

    dev = cuda_device
    W, H, R, thr, B = 160, 120, 32, 0.005, 4
    cam = Camera(W, H, W / 2, W / 2, W / 2, H / 2, pixel_center=0.5)
    hyp = syn.make_hypotheses(B, seed=0, device=dev)  # hypothesis 0 = the true pose
    grid = syn.sdf_mug(R, dev)[None].contiguous()
    cam_p = torch.tensor([[0.0, 0.0, 0.0], [0.12, 0.02, -0.05]], device=dev)
    cam_q = torch.nn.functional.normalize(torch.tensor([[0.0, 0.0, 0.0, 1.0], [0.02, 0.16, 0.01, 1.0]], device=dev), dim=1)
    p_c, q_c = views.to_camera_frames(hyp["position"][:1], hyp["orientation"][:1], cam_p, cam_q)
    obs = torch.cat([render_depth_batched(grid, p_c[v], q_c[v], hyp["inv_scale"][:1], thr, cam) for v in range(2)])
    assert float((obs[1] > 0).float().mean()) > 0.02  # the object is visible in the second view too
    kw = dict(sdf=grid, max_points=2000)
    two = HypothesisOptimizer(cam, thr, obs.contiguous(), hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                              camera_positions=cam_p, camera_orientations=cam_q, inlier_threshold=0.03, **kw)
    first = two.step().clone()
    assert bool(torch.isfinite(first).all())
    assert float(first[0]) < 0.5 * float(first[1:].mean())  # the true pose explains both views
    for _ in range(14):
        last = two.step()
    assert float(last[1:].mean()) < float(first[1:].mean())
    assert float(two.inlier_ratio[0]) > 0.9
    one = HypothesisOptimizer(cam, thr, obs[:1].contiguous(), hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                              camera_positions=cam_p[:1], camera_orientations=cam_q[:1], **kw)
    ref = HypothesisOptimizer(cam, thr, obs[0].contiguous(), hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                              optimizer="torch", **kw)
    for _ in range(3):
        torch.testing.assert_close(one.step(), ref.step(), rtol=1e-4, atol=1e-6)
    assert float((one.position.detach() - ref.position.detach()).abs().max()) < 1e-4


def _ptr(t):
    return None if t is None else t.data_ptr()


@pytest.mark.gpu
def test_view_kernels_match_the_host_mathematics(cuda_device):
    """sdfr_view_poses = views.to_camera_frames (the reference's rigid maps, golden-pinned above);
    sdfr_views_pull_back = views.pull_back plus the scale chain rule and the view loop's loss sum, and it
    clears what it consumed."""
    from sdfest_b200 import _lib

    dev, lib = cuda_device, _lib.lib()
    V, B = 3, 37
    g = torch.Generator().manual_seed(5)
    pos = torch.randn(B, 3, generator=g).to(dev)
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=1).to(dev)
    scale = (0.2 + torch.rand(B, generator=g)).to(dev)
    inv = (1.0 / scale).contiguous()
    cam_p = torch.randn(V, 3, generator=g).to(dev)
    cam_q = torch.nn.functional.normalize(torch.randn(V, 4, generator=g), dim=1).to(dev)
    pc, qc, isc = (torch.empty(V, B, 3, device=dev), torch.empty(V, B, 4, device=dev), torch.empty(V, B, device=dev))
    _lib.check(lib.sdfr_view_poses(_ptr(pos), _ptr(q), _ptr(inv), _ptr(cam_p), _ptr(cam_q), V, B, _ptr(pc),
                                   _ptr(qc), _ptr(isc), None), "view poses")
    want_p, want_q = views.to_camera_frames(pos.double(), q.double(), cam_p.double(), cam_q.double())
    torch.testing.assert_close(pc.double(), want_p, rtol=0, atol=2e-6)
    torch.testing.assert_close(qc.double(), want_q, rtol=0, atol=1e-6)
    assert torch.equal(isc, inv[None].expand(V, B))

    ins = {k: torch.randn(V, B, n, generator=g).to(dev).contiguous()
           for k, n in (("gr_p", 3), ("gr_q", 4), ("gr_is", 1), ("g2_p", 3), ("g2_q", 4), ("g2_s", 1), ("pl", 1))}
    ins["loss_sum"] = (10 * torch.rand(V, B, 1, generator=g)).to(dev).contiguous()
    ins["n"] = torch.randint(0, 4, (V, B, 1), generator=g).float().to(dev).contiguous()  # some views without overlap
    keep = {k: v.clone() for k, v in ins.items()}
    out = [torch.full((B, 3), 7.0, device=dev), torch.full((B, 4), 7.0, device=dev), torch.full((B,), 7.0, device=dev)]
    loss = torch.full((B,), 0.25, device=dev)
    w_d = 1.5
    _lib.check(lib.sdfr_views_pull_back(
        _ptr(cam_q), _ptr(scale), V, B, _ptr(ins["gr_p"]), _ptr(ins["gr_q"]), _ptr(ins["gr_is"]), _ptr(ins["g2_p"]),
        _ptr(ins["g2_q"]), _ptr(ins["g2_s"]), _ptr(ins["loss_sum"]), _ptr(ins["n"]), w_d, _ptr(ins["pl"]),
        _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(loss), _lib.STEP_CLEAR_INPUTS, None), "pull back")
    k = {a: b.double() for a, b in keep.items()}
    g_p, g_q = views.pull_back(k["gr_p"] + k["g2_p"], k["gr_q"] + k["g2_q"], cam_q.double())
    torch.testing.assert_close(out[0].double(), g_p, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[1].double(), g_q, rtol=1e-5, atol=1e-5)
    g_s = (k["g2_s"] - k["gr_is"] / scale.double()[None, :, None] ** 2).sum(0)[:, 0]
    torch.testing.assert_close(out[2].double(), g_s, rtol=1e-5, atol=1e-5)
    n = k["n"]
    per_view = torch.where(n > 0, w_d * k["loss_sum"] / n.clamp(min=1), torch.full_like(n, float("nan"))) + k["pl"]
    want_loss = 0.25 + per_view.sum(0)[:, 0]
    assert bool(torch.isnan(want_loss).any()) and not bool(torch.isnan(want_loss).all())
    torch.testing.assert_close(loss.double(), want_loss, rtol=1e-5, atol=1e-5, equal_nan=True)
    for v in ins.values():
        assert float(v.abs().max()) == 0.0  # consumed and cleared
    # arguments
    assert lib.sdfr_view_poses(None, None, None, None, None, 1, 1, None, None, None, None) == -1
    assert lib.sdfr_view_poses(None, None, None, None, None, 0, 5, None, None, None, None) == 0
    assert lib.sdfr_views_pull_back(None, None, 1, 1, *[None] * 8, 1.0, None, None, None, None, None, 0, None) == -1
    assert lib.sdfr_views_pull_back(None, None, -1, 1, *[None] * 8, 1.0, None, None, None, None, None, 0, None) == -2


@pytest.mark.gpu
def test_point_constraint_kernel_matches_the_reference_function_and_autograd(cuda_device):
    """sdfr_point_constraint against views.point_constraint_loss (pinned to the reference's own
    losses.point_constraint_loss by the golden test above), value and gradient, on UN-NORMALISED
    quaternions; both outputs accumulate."""
    from sdfest_b200 import _lib
    import ctypes

    dev, lib = cuda_device, _lib.lib()
    B = 50
    g = torch.Generator().manual_seed(9)
    q = (torch.randn(B, 4, generator=g) * (0.5 + torch.rand(B, 1, generator=g))).to(dev)
    src, tgt, weight = [0.3, -1.0, 0.2], [0.1, 0.4, -0.9], 2.5
    qd = q.double().requires_grad_(True)
    want = weight * views.point_constraint_loss(qd, torch.tensor(src, dtype=torch.float64, device=dev),
                                                torch.tensor(tgt, dtype=torch.float64, device=dev))
    want.sum().backward()
    g_raw = torch.full((B, 4), 0.5, device=dev)
    loss = torch.full((B,), 1.0, device=dev)
    _lib.check(lib.sdfr_point_constraint(_ptr(q), B, (ctypes.c_float * 3)(*src), (ctypes.c_float * 3)(*tgt), weight,
                                         _ptr(g_raw), _ptr(loss), None), "constraint")
    torch.testing.assert_close(loss.double() - 1.0, want.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(g_raw.double() - 0.5, qd.grad, rtol=1e-4, atol=1e-5)
    # exactly on target: the norm's subgradient is 0 (torch.linalg.norm), not NaN
    ident = torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=dev)
    g1, l1 = torch.zeros(1, 4, device=dev), torch.zeros(1, device=dev)
    _lib.check(lib.sdfr_point_constraint(_ptr(ident), 1, (ctypes.c_float * 3)(*src), (ctypes.c_float * 3)(*src), 1.0,
                                         _ptr(g1), _ptr(l1), None), "constraint")
    assert float(l1) == 0.0 and float(g1.abs().max()) == 0.0
    assert lib.sdfr_point_constraint(None, 1, None, None, 1.0, None, None, None) == -1
    assert lib.sdfr_point_constraint(None, 0, None, None, 1.0, None, None, None) == 0


def _view_scene(dev, B, R, with_decoder, V=2):
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.differentiable_renderer import Camera, render_depth_batched

    W, H, thr = 160, 120, 0.005
    cam = Camera(W, H, W / 2, W / 2, W / 2, H / 2, pixel_center=0.5)
    hyp = syn.make_hypotheses(B, seed=0, device=dev)
    grid = syn.sdf_mug(R, dev)[None].contiguous()
    cam_p = torch.tensor([[0.0, 0.0, 0.0], [0.12, 0.02, -0.05], [-0.1, 0.03, 0.02]], device=dev)[:V].contiguous()
    cam_q = torch.nn.functional.normalize(torch.tensor(
        [[0.0, 0.0, 0.0, 1.0], [0.02, 0.16, 0.01, 1.0], [0.05, -0.12, 0.02, 1.0]], device=dev), dim=1)[:V].contiguous()
    p_c, q_c = views.to_camera_frames(hyp["position"][:1], hyp["orientation"][:1], cam_p, cam_q)
    obs = torch.cat([render_depth_batched(grid, p_c[v], q_c[v], hyp["inv_scale"][:1], thr, cam)
                     for v in range(V)]).contiguous()
    # hypothesis 0 was the observed pose itself: at an exact optimum the gradient is rounding noise and Adam's
    # first steps (+-lr whatever the magnitude) amplify it, so two correct implementations part ways there
    hyp["position"] = (hyp["position"] + torch.tensor([0.004, -0.003, 0.005], device=dev)).contiguous()
    torch.manual_seed(0)
    if with_decoder:
        kw = dict(latent=0.1 * torch.randn(B, 8, device=dev), decoder=syn.residual_decoder(R, dev, syn.sdf_mug(R, dev)))
    else:
        kw = dict(sdf=grid)
    return cam, thr, obs, hyp, cam_p, cam_q, kw


@pytest.mark.gpu
@pytest.mark.parametrize("with_decoder", [False, True])
@pytest.mark.parametrize("constraint", [False, True])
def test_fused_views_match_the_autograd_composition(cuda_device, with_decoder, constraint):
    """The V-view iteration as C-ABI launches (view poses, per-view render / compare / point loss,
    pull-back, point constraint, step kernel) against the same iteration composed from the package's
    autograd operators, torch.optim.Adam and autograd through the rigid maps -- the structure of the
    reference's loop (simple_setup.py:408-463): losses step by step, parameters, inlier ratio."""
    from sdfest_b200.estimation import HypothesisOptimizer

    dev = cuda_device
    B, steps = 5, 4
    R = 64 if with_decoder else 32
    cam, thr, obs, hyp, cam_p, cam_q, kw = _view_scene(dev, B, R, with_decoder, V=3)
    extra = dict(point_constraint=(torch.tensor([0.0, 1.0, 0.0]), torch.tensor([0.1, 0.9, 0.2]), 0.05)) if constraint else {}

    def make(optimizer):
        if with_decoder:
            torch.manual_seed(0)
        return HypothesisOptimizer(cam, thr, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                                   camera_positions=cam_p, camera_orientations=cam_q, inlier_threshold=0.03,
                                   max_points=1500, optimizer=optimizer, **kw, **extra)

    a, b = make("torch"), make("fused")
    assert a.optimizer_impl == "torch" and b.optimizer_impl == "fused"
    for _ in range(steps):
        la, lb = a.step().clone(), b.step().clone()
        torch.testing.assert_close(lb, la, rtol=2e-3, atol=2e-5, equal_nan=True)
        torch.testing.assert_close(b.inlier_ratio, a.inlier_ratio, rtol=0, atol=2e-3)
    for name, lr in (("position", 1e-3), ("orientation", 1e-2), ("scale", 1e-3)):
        pa, pb = getattr(a, name).detach(), getattr(b, name).detach()
        assert float((pa - pb).abs().max()) < 0.05 * lr * steps, name
    if with_decoder:
        assert float((a.latent.detach() - b.latent.detach()).abs().max()) < 0.05 * 1e-2 * steps
    torch.testing.assert_close(torch.linalg.norm(b.orientation, dim=1), torch.ones(B, device=dev), rtol=0, atol=1e-6)
    torch.testing.assert_close(b.best_inlier_ratio, a.best_inlier_ratio, rtol=0, atol=2e-3)
    # graph replay continues the sequence
    c = make("fused")
    for _ in range(steps):
        c.step()
    b.capture(warmup=1)
    c.step(), c.step()
    lb = b.step().clone()
    torch.cuda.synchronize()
    torch.testing.assert_close(lb, c.last_losses, rtol=2e-3, atol=2e-5, equal_nan=True)


@pytest.mark.gpu
def test_fused_point_constraint_single_view(cuda_device):
    """One view, fixed grids: the constraint term on the fused path against the autograd composition."""
    from sdfest_b200.estimation import HypothesisOptimizer

    dev = cuda_device
    cam, thr, obs, hyp, _, _, kw = _view_scene(dev, 6, 32, False, V=1)
    con = (torch.tensor([0.0, 1.0, 0.0]), torch.tensor([0.0, 0.0, 1.0]), 2.0)

    def make(optimizer, **extra):
        return HypothesisOptimizer(cam, thr, obs[0].contiguous(), hyp["position"], hyp["orientation"],
                                   1.0 / hyp["inv_scale"], optimizer=optimizer, **kw, **extra)

    a, b, plain = make("torch", point_constraint=con), make("fused", point_constraint=con), make("fused")
    assert b.optimizer_impl == "fused"
    l_plain = plain.step().clone()
    for it in range(4):
        la, lb = a.step().clone(), b.step().clone()
        torch.testing.assert_close(lb, la, rtol=2e-3, atol=2e-5, equal_nan=True)
        if it == 0:
            want = 2.0 * views.point_constraint_loss(hyp["orientation"], con[0].to(dev), con[1].to(dev))
            torch.testing.assert_close(lb - l_plain, want, rtol=1e-4, atol=1e-5)
    assert float((a.orientation.detach() - b.orientation.detach()).abs().max()) < 0.05 * 1e-2 * 4
    assert float((b.orientation - plain.orientation).abs().max()) > 1e-4  # the constraint pulled on it
