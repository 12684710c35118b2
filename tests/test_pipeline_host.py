"""SDFPipeline (sdfest_b200/estimation/pipeline.py) on CPU: the reference's call sequence around the
loop -- masking in place, far field, initialisation from the first view's centred cloud, world-frame
conversion, the loop, result selection, errors -- with tiny stand-in networks and a differentiable
stand-in for the CUDA renderer (the loop's kernels are tested on the GPU in test_hypothesis_step.py)."""
import pytest
import torch
from torch import nn

from sdfest_b200.estimation import hypotheses, views
from sdfest_b200.estimation.pipeline import NoDepthError, SDFPipeline

W, H = 16, 12
CONFIG = {"device": "cpu", "camera": dict(width=W, height=H, fx=14.0, fy=14.0, cx=8.0, cy=6.0, pixel_center=0.5),
          "threshold": 0.005, "max_iterations": 6, "depth_weight": 1.0, "pc_weight": 3.0,
          "init": {"backbone_type": "VanillaPointNet", "normalize_pose": True, "head": {"orientation_repr": "quaternion"}}}


class TinyVAE(nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.decoder = nn.Linear(2, 8 ** 3)
        with torch.no_grad():
            self.decoder.weight.copy_(0.05 * torch.randn(8 ** 3, 2, generator=g))
            self.decoder.bias.copy_(torch.rand(8 ** 3, generator=g) - 0.3)

    def decode(self, z):
        return self.decoder(z).view(-1, 1, 8, 8, 8)


class TinyInit(nn.Module):
    """Returns a fixed estimate relative to the (centred) input cloud and records what it was given."""

    def __init__(self):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(0.3))
        self.seen = None

    def forward(self, x):
        self.seen = x.detach().clone()
        return (torch.tensor([[0.2, -0.1]]), torch.tensor([[0.01, 0.0, -0.02]]), self.scale.reshape(1) * 1.0,
                torch.tensor([[0.0, 0.0, 0.0, 1.0]]))


def _fake_render_and_compare(sdf, position, orientation, inv_scale, depth_obs, threshold, camera):
    B = position.shape[0]
    target = depth_obs[depth_obs > 0].mean()
    loss = (position[:, 2] + target) ** 2 + 0.1 * (orientation[:, 3] - 1) ** 2 + 0.01 * inv_scale + 0.0 * sdf.sum()
    return loss, depth_obs[None].expand(B, -1, -1) * 1.02, torch.ones(B)


@pytest.fixture
def scene(monkeypatch):
    monkeypatch.setattr(hypotheses, "render_and_compare", _fake_render_and_compare)
    from oracle.pc_loss import point_loss
    from sdfest_b200.estimation import losses

    monkeypatch.setattr(losses, "point_loss", point_loss)  # the product's point loss is CUDA-only
    depth = torch.full((H, W), 0.8)
    depth[0, 0] = 5.0  # far-field outlier inside the mask
    mask = torch.zeros(H, W, dtype=torch.bool)
    mask[3:9, 4:11] = True
    mask[0, 0] = True
    return depth, mask


def test_call_sequence_and_result(scene):
    depth, mask = scene
    init = TinyInit()
    pipe = SDFPipeline(dict(CONFIG, far_field=2.0), TinyVAE(), init)
    given = depth.clone()
    position, orientation, scale, latent = pipe(given, mask, None)
    # masking and the far field act in place on the caller's tensor, like the reference
    assert float(given[0, 0]) == 0.0 and float(given[~mask].abs().max()) == 0.0 and int((given > 0).sum()) == 42
    # the network saw the centred cloud of the first view
    assert init.seen.shape == (1, 42, 3) and float(init.seen.mean(1).abs().max()) < 1e-6
    assert position.shape == (1, 3) and orientation.shape == (1, 4) and scale.shape == (1,) and latent.shape == (1, 2)
    opt = pipe.last_optimizer
    assert opt.optimizer_impl == "torch" and int(opt._iteration[0]) == 6
    # the optimisation started from centroid + network offset and moved
    from sdfest_b200.estimation import depth_to_pointcloud
    centroid = depth_to_pointcloud(given, pipe.cam).mean(0)
    start = centroid + torch.tensor([0.01, 0.0, -0.02])
    assert 0 < float((position[0] - start).abs().max()) < 0.02
    assert float((latent - torch.tensor([[0.2, -0.1]])).abs().max()) > 0  # the shape was optimised
    torch.testing.assert_close(torch.linalg.norm(orientation, dim=1), torch.ones(1))
    assert float(opt.inlier_ratio[0]) == 1.0  # 2 % error everywhere


def test_options_and_errors(scene):
    depth, mask = scene
    vae, init = TinyVAE(), TinyInit()
    # no shape optimisation: the latent comes back untouched
    pipe = SDFPipeline(CONFIG, vae, init)
    _, _, _, latent = pipe(depth.clone(), mask, None, shape_optimization=False)
    assert torch.equal(latent, torch.tensor([[0.2, -0.1]]))
    # mean shape
    _, _, _, latent = SDFPipeline(dict(CONFIG, mean_shape=True), vae, init)(depth.clone(), mask, None, shape_optimization=False)
    assert float(latent.abs().max()) == 0.0
    # best inlier ratio returns the kept copy
    best = SDFPipeline(dict(CONFIG, result_selection_strategy="best_inlier_ratio"), vae, init)
    p_best = best(depth.clone(), mask, None)[0]
    assert torch.equal(p_best, best.last_optimizer.best_position)
    # a camera pose: the initial estimate is moved to the world frame, the loop runs per view
    cam_p = torch.tensor([0.1, 0.0, 0.05])
    cam_q = torch.nn.functional.normalize(torch.tensor([0.0, 0.2, 0.0, 1.0]), dim=0)
    posed = SDFPipeline(dict(CONFIG, max_iterations=1), vae, init)
    posed(depth.clone(), mask, None, camera_positions=cam_p, camera_orientations=cam_q)
    opt = posed.last_optimizer
    assert opt._views is not None and opt.depth_obs.shape == (1, H, W)
    p_c, _ = views.to_camera_frames(opt.best_position * 0 + opt.position.detach(), opt.orientation.detach(),
                                    cam_p[None], cam_q[None])
    # back in the camera frame the object sits at the cloud centroid + the network's offset (one Adam step away)
    z_mean = float(depth[mask].mean())
    assert abs(float(p_c[0, 0, 2]) + z_mean + 0.02) < 5e-3
    # several hypotheses (extension): the returned one is a single estimate
    multi = SDFPipeline(dict(CONFIG, n_hypotheses=4), vae, init)
    p = multi(depth.clone(), mask, None)[0]
    assert p.shape == (1, 3) and multi.last_optimizer.position.shape == (4, 3)
    # point constraint goes through
    SDFPipeline(dict(CONFIG, max_iterations=1), vae, init)(
        depth.clone(), mask, None, point_constraint=(torch.tensor([0.0, 1, 0]), torch.tensor([0.0, 0, 1]), 1.0))

    with pytest.raises(NoDepthError):
        pipe(depth.clone(), torch.zeros(H, W, dtype=torch.bool), None)
    for kw in (dict(visualize=True), dict(log_path="x"), dict(animation_path="x"),
               dict(prior_orientation_distribution=torch.ones(1, 4))):
        with pytest.raises(NotImplementedError):
            pipe(depth.clone(), mask, None, **kw)
    with pytest.raises(ValueError):
        SDFPipeline(dict(CONFIG, result_selection_strategy="median"), vae, init)
    with pytest.raises(NotImplementedError):
        SDFPipeline(dict(CONFIG, init_view="best"), vae, init)(depth.clone(), mask, None)
