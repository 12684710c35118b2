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
