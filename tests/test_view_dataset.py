"""Batched view generation (reference generated_dataset.py:21-342) and the
runtime-breakdown harness (8f rank 4; reference real_data.py:217-319)."""
import math

import pytest
import torch

from sdfest_b200.differentiable_renderer import Camera
from sdfest_b200.estimation import BatchedSDFViewGenerator
from sdfest_b200.estimation import view_dataset as vd

CFG = dict(width=64, height=48, z_min=0.5, z_max=1.0, extent_mean=0.2, extent_std=0.02)


class FakeRender:
    """Stands in for render_depth_batched on the CPU: a rectangle at the object's distance; the
    first call returns an empty view for sample 0 to exercise the re-draw."""

    def __init__(self):
        self.calls = 0

    def __call__(self, sdf, p, q, inv_s, thr, cam):
        d = torch.zeros(p.shape[0], cam.height, cam.width)
        d[:, 10:20, 10:30] = (-p[:, 2])[:, None, None]
        if self.calls == 0:
            d[0] = 0
        self.calls += 1
        return d


def make(cfg, **kw):
    r = FakeRender()
    g = BatchedSDFViewGenerator({**CFG, **cfg}, lambda z: torch.zeros(z.shape[0], 1, 8, 8, 8), 8, 4,
                                "cpu", render=r, **kw)
    return g, r


def test_defaults_and_required_keys_follow_the_reference():
    assert vd.DEFAULT_CONFIG["render_threshold"] == 0.004 and vd.DEFAULT_CONFIG["fov_deg"] == 90
    with pytest.raises(KeyError, match="z_min"):
        BatchedSDFViewGenerator({"z_max": 1.0, "extent_mean": 1, "extent_std": 0}, None, 8, 2, "cpu")
    with pytest.raises(NotImplementedError):
        make({"orientation_repr": "discretized"})
    g, _ = make({})
    f = 64 / math.tan(math.radians(90) / 2) / 2  # generated_dataset.py:134
    assert g.camera.fx == pytest.approx(f) and g.camera.cx == 32 and g.camera.pixel_center == 0.5


def test_pose_sampling_ranges():
    cam = Camera(640, 480, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
    cfg = {**vd.DEFAULT_CONFIG, **CFG}
    p, q, s = vd.sample_poses(4096, cam, cfg, torch.Generator().manual_seed(0), "cpu")
    z = -p[:, 2]
    assert float(z.min()) >= 0.5 and float(z.max()) <= 1.0
    x_pix, y_pix = p[:, 0] / z * 320.0, p[:, 1] / z * 320.0
    assert float(x_pix.min()) >= -320.001 and float(x_pix.max()) <= 240.001  # sic, :264
    assert float(y_pix.min()) >= -240.001 and float(y_pix.max()) <= 240.001
    assert torch.allclose(q.norm(dim=1), torch.ones(4096), atol=1e-5)
    assert abs(float(s.mean()) - 0.1) < 2e-3 and abs(float(s.std()) - 0.01) < 2e-3


def test_empty_views_are_redrawn_and_batches_are_reproducible():
    g, r = make({})
    s = g.generate()
    assert r.calls == 2 and float(s["depth"].flatten(1).amax(1).min()) > 0
    assert set(s) == {"depth", "latent_shape", "position", "quaternion", "orientation", "scale"}
    g2, _ = make({})
    s2 = g2.generate()
    assert torch.equal(s["position"], s2["position"]) and torch.equal(s["latent_shape"], s2["latent_shape"])


def test_noise_models_and_pointsets():
    k = vd.gaussian_kernel(5, 1.0, "cpu")
    assert k.shape == (1, 1, 5, 5) and float(k.sum()) == pytest.approx(1.0) and float(k[0, 0, 2, 2]) == float(k.max())
    g, _ = make({"mask_noise": True, "gaussian_noise_probability": 1.0, "pointcloud": True,
                 "normalize_pose": True, "scale_to_unit_ball": True})
    s = g.generate()
    d = s["depth"]
    assert not torch.isnan(d).any() and float(d.min()) >= 0
    assert float(d[:, :5].abs().max()) == 0  # far from the mask: stays background
    for pts in s["pointset"]:
        assert pts.shape[1] == 3 and pts.shape[0] > 0
        assert float(torch.linalg.norm(pts)) == pytest.approx(1.0, rel=1e-4)
        assert float(pts.mean(0).abs().max()) < 1e-3
    # identity transform of the mask perturbation leaves masks unchanged in the interior
    m = torch.zeros(2, 48, 64, dtype=torch.bool)
    m[:, 10:30, 20:40] = True
    pm = g.perturb_masks(m)
    assert pm.shape == m.shape and pm.dtype == torch.bool and bool(pm[:, 15:25, 25:35].all())
    assert int((pm != m).sum()) < 200


@pytest.mark.gpu
def test_generator_renders_real_views_on_gpu(cuda_device):
    from sdfest_b200 import synthetic as syn

    dec = syn.residual_decoder(64, cuda_device, syn.sdf_mug(64, cuda_device))
    cfg = dict(width=160, height=120, z_min=0.4, z_max=0.8, extent_mean=0.25, extent_std=0.02,
               pointcloud=True, gaussian_noise_probability=0.5)
    g = BatchedSDFViewGenerator(cfg, dec, 8, 8, cuda_device, seed=1)
    s = g.generate()
    d = s["depth"]
    assert d.shape == (8, 120, 160) and float(d.flatten(1).amax(1).min()) > 0
    z, r = -s["position"][:, 2], s["scale"] * math.sqrt(3.0)
    hit = d > 0
    zmin = torch.where(hit, d, torch.full_like(d, 1e9)).flatten(1).amin(1)
    assert bool((zmin > z - r - 0.05).all()) and bool((d.flatten(1).amax(1) < z + r + 0.05).all())
    assert len(s["pointset"]) == 8 and all(int(p.shape[0]) == int(h.sum()) for p, h in zip(s["pointset"], hit))


@pytest.mark.gpu
def test_runtime_breakdown_reports_every_phase(cuda_device):
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.differentiable_renderer import render_depth_batched
    from sdfest_b200.estimation import HypothesisOptimizer, runtime_analysis

    dev = cuda_device
    B, R, thr = 4, 64, 0.005
    cam = Camera(160, 120, 80.0, 80.0, 80.0, 60.0, pixel_center=0.5)
    hyp = syn.make_hypotheses(B, seed=5, device=dev)
    base = syn.sdf_mug(R, dev)
    obs = render_depth_batched(base, hyp["position"][:1], hyp["orientation"][:1], hyp["inv_scale"][:1],
                               thr, cam)[0].contiguous()
    def make(optimizer):
        return HypothesisOptimizer(cam, thr, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                                   latent=torch.zeros(B, 8, device=dev),
                                   decoder=syn.residual_decoder(R, dev, base), optimizer=optimizer)

    res = runtime_analysis.phase_breakdown(make("torch"), iterations=3, warmup=1)
    assert set(res) == set(runtime_analysis.PHASES) | {"total"} and all(v > 0 for v in res.values())
    with pytest.raises(ValueError):
        runtime_analysis.phase_breakdown(make("fused"), iterations=1, warmup=0)
    for optimizer in ("torch", "fused"):
        assert runtime_analysis.iteration_ms(make(optimizer), iterations=3, warmup=1, graph=True) > 0


@pytest.mark.gpu
def test_streamed_decode_render_compare_matches_the_operator(cuda_device):
    """Host latents / poses in, losses and gradients out through one CUDA graph
    (estimation.StreamedDecodeRenderCompare) == decode_render_compare + autograd on device tensors; a second set
    of host buffers re-captures; eager mode agrees with the replay."""
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.differentiable_renderer import Camera, render_depth_batched
    from sdfest_b200.estimation import StreamedDecodeRenderCompare, decode_render_compare

    dev = cuda_device
    B, R, W, H, thr = 5, 64, 160, 120, 0.005
    cam = Camera(W, H, W / 2, W / 2, W / 2, H / 2, pixel_center=0.5)
    dec = syn.residual_decoder(R, dev, syn.sdf_mug(R, dev))
    hyp = syn.make_hypotheses(B, seed=0, device=dev)
    base = syn.make_hypotheses(1, seed=0, device=dev)
    obs = render_depth_batched(syn.hypothesis_grids(base["shape_param"], R, dev), base["position"],
                               base["orientation"], base["inv_scale"], thr, cam)[0].contiguous()
    s = StreamedDecodeRenderCompare(dec, cam, thr, B, 8, dev)
    assert s.h2d_bytes == 4 * (B * 16 + W * H) and s.d2h_bytes == 4 * B * 18
    for seed in (0, 1):
        lat = 0.2 * torch.randn(B, 8, generator=torch.Generator().manual_seed(seed))
        host = [t.cpu().pin_memory() for t in (lat, hyp["position"] + 0.002 * seed, hyp["orientation"],
                                               1.0 / hyp["inv_scale"], obs)]
        out = {k: (v.clone() if v is not None else None) for k, v in s(*host).items()}
        again = s(*host)  # replay
        leaves = [t.to(dev).requires_grad_(True) for t in host[:4]]
        w, b = dec.tail_parameters()
        loss, depth, n, _ = decode_render_compare(dec.trunk(leaves[0]), w, b, leaves[1], leaves[2], leaves[3], obs, None,
                                                  R, thr, cam, base=dec.base, depth_weight=1.0, pc_weight=0.0)
        g = torch.autograd.grad(loss.sum(), leaves)
        for res in (out, again, s(*host, graph=False)):
            torch.testing.assert_close(res["loss"], loss.detach().cpu(), rtol=1e-4, atol=1e-6)
            assert torch.equal(res["n_overlap"], n.cpu())
            for key, want in (("g_latent", g[0]), ("g_position", g[1]), ("g_orientation", g[2]), ("g_scale", g[3])):
                want = want.cpu()
                assert float((res[key] - want).abs().max()) <= 2e-3 * float(want.abs().max()), key
        assert torch.equal(s.depth, depth)
    with pytest.raises(TypeError):
        StreamedDecodeRenderCompare(dec.decoder, cam, thr, B, 8, dev)
