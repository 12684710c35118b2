"""Drop-in proof: the reference's OWN callers of the renderer, unmodified, on top of this library.

The reference package (installed as-is into baseline/_ref by oracle/install_reference.py) is imported
twice, bound once to ``sdfest_b200.compat.sdf_renderer_cpp`` (libsdfrender.so behind the reference's
pybind interface, replacing only the JIT ``load`` of sdf_renderer.py:21-28) and once to the reference's
own CUDA extension (oracle/_ref).  Then the reference's callers run on both and are compared:

 * the optimisation loop of ``differentiable_renderer/scripts/experiments.py:103-129`` (Adam on pose and
   inverse scale against a rendered reference image) -- loss trajectories;
 * ``estimation/simple_setup.py::SDFPipeline.__call__`` (:213-600) with the reference's trained mug VAE
   (its test fixture mug.pt), 30 iterations -- returned position / orientation / scale / latent;
 * this package's ``sdfest_b200.estimation.SDFPipeline`` (fused CUDA-graph loop) on the same networks and
   observation against the reference pipeline's result;
 * ``initialization/datasets/generated_dataset.py::SDFVAEViewDataset`` (the third caller) -- one sample.

The CPU part (no GPU needed) checks that the reference imports on the shim and that the shim has the
pybind module's interface.
"""
import importlib
import inspect
import os
import re

import numpy as np
import pytest
import torch

import ref_loader
from util import mug_sdf, shoemake

needs_reference = pytest.mark.skipif(not ref_loader.available(),
                                     reason="baseline/_ref not installed (python oracle/install_reference.py)")
CAMERA = dict(width=640, height=480, fx=320, fy=320, cx=320, cy=240, pixel_center=0.5)  # default.yaml:1-8


def _shim():
    from sdfest_b200.compat import sdf_renderer_cpp

    return sdf_renderer_cpp


@needs_reference
def test_reference_imports_on_the_shim_without_a_gpu():
    shim = _shim()
    ref_loader.load_reference(shim)
    renderer = importlib.import_module("sdfest.differentiable_renderer.sdf_renderer")
    assert renderer.sdf_renderer_cpp is shim
    # the two call sites of the native module are the reference's own, untouched (py:311, py:347)
    src = inspect.getsource(renderer.SDFRendererFunctionGPU)
    assert "sdf_renderer_cpp.forward(" in src and "sdf_renderer_cpp.backward(" in src
    assert list(inspect.signature(shim.forward).parameters) == [
        "sdf", "position", "orientation", "inv_scale", "width", "height", "cx", "cy", "fx", "fy", "threshold"]
    assert list(inspect.signature(shim.backward).parameters) == [
        "grad_depth_image", "depth_image", "sdf", "position", "orientation", "inv_scale", "width", "height",
        "cx", "cy", "fx", "fy"]
    setup = importlib.import_module("sdfest.estimation.simple_setup")
    assert setup.render_depth_gpu is renderer.render_depth_gpu
    with pytest.raises(RuntimeError, match="CUDA tensor"):  # CHECK_CUDA of sdf_renderer.cpp:9
        shim.forward(torch.zeros(4, 4, 4), torch.zeros(3), torch.zeros(4), torch.ones(1), 8, 8, 4.0, 4.0, 4.0,
                     4.0, 0.01)
    ref_loader.purge()


# ------------------------------------------------------------------------------------------
# GPU: the reference's callers on both native modules
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_ext():
    from oracle import build_ref

    mod = build_ref.load_module()
    if mod is None:
        pytest.skip("oracle/_ref/sdf_renderer_cpp.so not present")
    return mod


def _run_experiment(native, sdf_path, capsys, steps):
    ref_loader.load_reference(native)
    ex = importlib.import_module("sdfest.differentiable_renderer.scripts.experiments")
    capsys.readouterr()
    # the click command's undecorated function: the reference's loop, as shipped (experiments.py:33-155)
    ex.offset_experiment.callback(
        sdf_path=sdf_path, ref_sdf_path=None, pos=(0.02, -0.01, -0.4), rot=(400.0, 90.0, 0.0), scale=0.15,
        pos_off=(0.01, -0.01, 0.01), rot_off=(4.0, -3.0, 2.0), scale_off=0.01, steps=steps, width=640, height=480,
        fov=90.0, gpu=True, visualize=None, threshold=0.01)
    out = capsys.readouterr().out
    ref_loader.purge()
    return np.array([float(m) for m in re.findall(r"loss: ([0-9.eE+-]+) step", out)])


@pytest.mark.gpu
@needs_reference
def test_reference_experiments_loop_on_both_native_modules(cuda_device, ref_ext, tmp_path, capsys):
    sdf_path = str(tmp_path / "mug.npy")
    np.save(sdf_path, mug_sdf())
    steps = 25
    ours = _run_experiment(_shim(), sdf_path, capsys, steps)
    theirs = _run_experiment(ref_ext, sdf_path, capsys, steps)
    assert len(ours) == steps and len(theirs) == steps
    assert ours[-1] < 0.7 * ours[0], "the loop must actually optimise"
    # same trajectory: fp32 atomics reorder the gradient sums, Adam amplifies that slowly
    np.testing.assert_allclose(ours, theirs, rtol=2e-3)
    np.testing.assert_allclose(ours[:3], theirs[:3], rtol=1e-5)


def _pipeline_config(init_path, vae_yaml, vae_path, iterations):
    import yaml

    vae_cfg = yaml.safe_load(open(vae_yaml))
    return {
        "device": "cuda", "camera": dict(CAMERA), "threshold": 0.005, "far_field": None,
        "max_iterations": iterations, "depth_weight": 1.0, "pc_weight": 3.0, "nn_weight": 0.0,
        "mean_shape": False, "init_view": "first", "shape_init": "prediction", "iso_threshold": 0.02,
        "init": {"backbone_type": "VanillaPointNet",
                 "backbone": {"in_size": 3, "mlp_out_sizes": [128, 128, 128, 128, 1024], "batchnorm": True,
                              "dense": True, "residual": True},
                 "head_type": "SDFPoseHead",
                 "head": {"in_size": 1024, "mlp_out_sizes": [512, 256, 128], "batchnorm": True,
                          "orientation_repr": "quaternion"},
                 "normalize_pose": True, "model": init_path},
        "vae": {"latent_size": vae_cfg["latent_size"], "encoder": vae_cfg["encoder"], "decoder": vae_cfg["decoder"],
                "model": vae_path},
    }


def _make_init_weights(setup, cfg, true_q, path):
    """A randomly initialised SDFPoseNet of the reference's architecture whose last layer is overwritten to
    emit a fixed, plausible initial estimate (there are no trained initialisation weights offline): latent
    0.3 * N(0,1), position = cloud centroid + (0.01, -0.01, -0.04), scale 0.14, orientation 8 degrees off."""
    torch.manual_seed(5)
    net = setup.SDFPoseNet(
        setup.INIT_MODULE_DICT["VanillaPointNet"](**cfg["init"]["backbone"]),
        setup.INIT_MODULE_DICT["SDFPoseHead"](shape_dimension=cfg["vae"]["latent_size"], **cfg["init"]["head"]))
    L = cfg["vae"]["latent_size"]
    ang = np.deg2rad(8.0)
    dq = np.array([np.sin(ang / 2), 0.0, 0.0, np.cos(ang / 2)])
    x1, y1, z1, w1 = dq
    x2, y2, z2, w2 = true_q
    q0 = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                   w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    with torch.no_grad():
        net._head._final_layer.weight.zero_()
        net._head._final_layer.bias.copy_(torch.cat([
            0.3 * torch.randn(L), torch.tensor([0.01, -0.01, -0.04]), torch.tensor([0.14]),
            torch.as_tensor(q0, dtype=torch.float32)]))
    torch.save(net.state_dict(), path)


def _observation(dev, vae_path, vae_yaml):
    """A masked depth image of the reference's trained mug VAE at a known pose, rendered by this library."""
    from sdfest_b200.differentiable_renderer import Camera, render_depth_gpu

    ref_loader.load_reference(_shim())
    vae_mod = importlib.import_module("sdfest.vae.sdf_vae")
    import yaml

    c = yaml.safe_load(open(vae_yaml))
    vae = vae_mod.SDFVAE(sdf_size=64, latent_size=c["latent_size"], encoder_dict=c["encoder"], decoder_dict=c["decoder"],
                         device=dev).to(dev)
    vae.load_state_dict(torch.load(vae_path, map_location=dev))
    vae.eval()
    z_true = torch.tensor([[0.4, -0.3, 0.2, 0.0, -0.5, 0.3, 0.1, -0.2]], device=dev)
    with torch.no_grad():
        sdf = vae.decode(z_true)[0, 0].contiguous()
    q_true = shoemake(1)
    depth = render_depth_gpu(sdf, torch.tensor([0.02, -0.01, -0.45], device=dev), torch.as_tensor(q_true, device=dev),
                             torch.tensor([1 / 0.15], device=dev), threshold=0.005, camera=Camera(**CAMERA))
    ref_loader.purge()
    return depth, q_true


def _run_reference_pipeline(native, cfg, depth, make_init=None):
    ref_loader.load_reference(native)
    setup = importlib.import_module("sdfest.estimation.simple_setup")
    if make_init is not None:
        make_init(setup)
    pipe = setup.SDFPipeline(cfg)
    d = depth.clone()
    out = pipe(d, d > 0, torch.zeros(*d.shape, 3, device=d.device))
    out = tuple(t.detach().clone() for t in out)
    nets = (pipe.vae, pipe.init_network)
    ref_loader.purge()
    return out, nets


@pytest.mark.gpu
@needs_reference
def test_reference_pipeline_on_both_native_modules_and_ours(cuda_device, ref_ext, tmp_path):
    dev = cuda_device
    vae_path, vae_yaml = os.path.join(ref_loader.FIXTURES, "mug.pt"), os.path.join(ref_loader.FIXTURES, "mug.yaml")
    init_path = str(tmp_path / "init.pt")
    depth, q_true = _observation(dev, vae_path, vae_yaml)
    assert int((depth > 0).sum()) > 5000
    iterations = 30
    cfg = _pipeline_config(init_path, vae_yaml, vae_path, iterations)
    cudnn = torch.backends.cudnn.enabled
    try:
        ours, nets = _run_reference_pipeline(_shim(), cfg, depth,
                                             make_init=lambda setup: _make_init_weights(setup, cfg, q_true, init_path))
        theirs, _ = _run_reference_pipeline(ref_ext, cfg, depth)
    finally:
        torch.backends.cudnn.enabled = cudnn  # the reference pipeline switches cuDNN off globally (setup:46)
    names = ("position", "orientation", "scale", "latent")
    tol = dict(position=2e-3, orientation=1e-2, scale=2e-3, latent=5e-2)  # absolute; 30 chaotic Adam steps
    for a, b, nm in zip(ours, theirs, names):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= tol[nm], (nm, a, b)
    assert torch.isfinite(torch.cat([t.flatten() for t in ours])).all()

    # this package's pipeline (fused iteration, CUDA-graph replay) with the SAME networks and observation
    from sdfest_b200.estimation import SDFPipeline

    vae, init_network = nets
    mine = SDFPipeline(dict(cfg, relative_inlier_threshold=0.03), vae, init_network)
    d = depth.clone()
    got = mine(d, d > 0, None)
    assert mine.last_optimizer.optimizer_impl == "fused"
    for a, b, nm in zip(got, ours, names):
        assert tuple(a.shape) == tuple(b.shape), (nm, a.shape, b.shape)
        assert float((a - b).abs().max()) <= tol[nm], (nm, a, b)


@pytest.mark.gpu
@needs_reference
def test_pipeline_with_views_and_point_constraint_matches_the_reference(cuda_device, ref_ext, tmp_path):
    """Two views with camera poses and a point constraint (simple_setup.py:164-175, 420-446): the reference's
    pipeline on ITS OWN extension against this package's SDFPipeline, where all of it runs on the fused
    path (sdfr_view_poses, per-view render / compare / point loss, sdfr_views_pull_back,
    sdfr_point_constraint, one sdfr_hypothesis_step) from a CUDA graph."""
    from sdfest_b200.differentiable_renderer import Camera, render_depth_gpu
    from sdfest_b200.estimation import SDFPipeline, views

    dev = cuda_device
    vae_path, vae_yaml = os.path.join(ref_loader.FIXTURES, "mug.pt"), os.path.join(ref_loader.FIXTURES, "mug.yaml")
    init_path = str(tmp_path / "init.pt")
    depth0, q_true = _observation(dev, vae_path, vae_yaml)
    iterations = 20
    cfg = _pipeline_config(init_path, vae_yaml, vae_path, iterations)
    cam_p = torch.tensor([[0.0, 0.0, 0.0], [0.10, 0.02, -0.04]], device=dev)
    cam_q = torch.nn.functional.normalize(torch.tensor([[0.0, 0.0, 0.0, 1.0], [0.02, 0.14, 0.01, 1.0]], device=dev), dim=1)
    constraint = (torch.tensor([0.0, 1.0, 0.0], device=dev), torch.tensor([0.05, 0.95, 0.1], device=dev), 0.02)

    ref_loader.load_reference(ref_ext)
    cudnn = torch.backends.cudnn.enabled
    try:
        setup = importlib.import_module("sdfest.estimation.simple_setup")
        _make_init_weights(setup, cfg, q_true, init_path)
        pipe = setup.SDFPipeline(cfg)
        # the second view: the same object (true latent / pose of _observation) seen from camera 1
        z_true = torch.tensor([[0.4, -0.3, 0.2, 0.0, -0.5, 0.3, 0.1, -0.2]], device=dev)
        with torch.no_grad():
            sdf = pipe.vae.decode(z_true)[0, 0].contiguous()
        p_c, q_c = views.to_camera_frames(torch.tensor([[0.02, -0.01, -0.45]], device=dev),
                                          torch.as_tensor(q_true, device=dev, dtype=torch.float32)[None], cam_p, cam_q)
        depth1 = render_depth_gpu(sdf, p_c[1, 0].contiguous(), q_c[1, 0].contiguous(), torch.tensor([1 / 0.15], device=dev),
                                  threshold=0.005, camera=Camera(**CAMERA))
        assert int((depth1 > 0).sum()) > 3000
        depths = torch.stack([depth0, depth1]).contiguous()
        d = depths.clone()
        theirs = pipe(d, d > 0, torch.zeros(*d.shape, 3, device=dev), camera_positions=cam_p.clone(),
                      camera_orientations=cam_q.clone(), point_constraint=constraint)
        theirs = tuple(t.detach().clone() for t in theirs)
        vae, init_network = pipe.vae, pipe.init_network
    finally:
        torch.backends.cudnn.enabled = cudnn
        ref_loader.purge()

    mine = SDFPipeline(dict(cfg, relative_inlier_threshold=0.03), vae, init_network)
    d = depths.clone()
    got = mine(d, d > 0, None, camera_positions=cam_p, camera_orientations=cam_q, point_constraint=constraint)
    opt = mine.last_optimizer
    assert opt.optimizer_impl == "fused" and opt._V == 2 and opt._graph is not None
    tol = dict(position=2e-3, orientation=1e-2, scale=2e-3, latent=5e-2)  # absolute; chaotic Adam steps
    for a, b, nm in zip(got, theirs, ("position", "orientation", "scale", "latent")):
        assert tuple(a.shape) == tuple(b.shape), (nm, a.shape, b.shape)
        assert float((a - b).abs().max()) <= tol[nm], (nm, a, b)


@pytest.mark.gpu
@needs_reference
def test_reference_view_dataset_on_the_shim(cuda_device):
    """initialization/datasets/generated_dataset.py:277-284 calls render_depth_gpu with keyword arguments
    and no gradient; one sample through the unmodified class."""
    vae_path, vae_yaml = os.path.join(ref_loader.FIXTURES, "mug.pt"), os.path.join(ref_loader.FIXTURES, "mug.yaml")
    import yaml

    c = yaml.safe_load(open(vae_yaml))
    ref_loader.load_reference(_shim())
    try:
        gd = importlib.import_module("sdfest.initialization.datasets.generated_dataset")
        vae_mod = importlib.import_module("sdfest.vae.sdf_vae")
        vae = vae_mod.SDFVAE(sdf_size=64, latent_size=c["latent_size"], encoder_dict=c["encoder"],
                             decoder_dict=c["decoder"], device=cuda_device).to(cuda_device)
        vae.load_state_dict(torch.load(vae_path, map_location=cuda_device))
        cfg = gd.SDFVAEViewDataset.default_config if hasattr(gd.SDFVAEViewDataset, "default_config") else {}
        config = dict(cfg)
        config.update(dict(width=640, height=480, fov_deg=90, z_min=0.3, z_max=0.8, extent_mean=0.2, extent_std=0.03,
                           pointcloud=False, render_threshold=0.005, orientation_repr="quaternion",
                           mask_noise=False, norm_noise=False, scale_to_unit_ball=False, camera=dict(CAMERA),
                           normalize_pose=False))
        torch.manual_seed(0)
        ds = gd.SDFVAEViewDataset(config=config, vae=vae)
        sample = next(iter(ds))
        depth = sample["depth"] if "depth" in sample else sample["pointset"]
        assert tuple(depth.shape) == (480, 640) and float((depth > 0).float().mean()) > 0.001
    finally:
        ref_loader.purge()
