"""The oracle against the golden vectors produced by the REFERENCE's own CPU renderer
(tests/golden/make_golden.py imports sdfest/differentiable_renderer/simple_renderer.py).
CPU-only; this is what pins the oracle."""
import numpy as np
import pytest

import oracle
from util import golden_names, load_golden

NAMES = golden_names()


def test_golden_set_present():
    assert len(NAMES) >= 6


@pytest.mark.parametrize("name", NAMES)
def test_f64_oracle_reproduces_reference_forward(name):
    z = load_golden(name)
    d, steps, t = oracle.render(z["sdf"], z["position"], z["orientation"], z["inv_scale"],
                                z["W"], z["H"], threshold=float(z["threshold"]),
                                dtype=np.float64, extras=True, **z["cam"])
    gold = z["depth"]
    assert ((gold > 0) == (d > 0)).all()
    hit = gold > 0
    assert np.abs(d - gold)[hit].max() / gold[hit].max() < 1e-12
    # value="c" of the reference: number of trilinear samples of every ray that hit
    assert (np.where(hit, steps, 0) == z["steps"]).all()
    # depth = |t * d_z| (simple_renderer.py:304)
    assert (t[hit] > 0).all() and (d[hit] <= t[hit] * (1 + 1e-12)).all()


@pytest.mark.parametrize("name", NAMES)
def test_f64_oracle_reproduces_reference_derivatives(name):
    z = load_golden(name)
    bw = oracle.render_backward(z["g"], z["depth"], z["sdf"], z["position"], z["orientation"],
                                z["inv_scale"], z["W"], z["H"], sdf_grad_mode="exact",
                                dtype=np.float64, want_deriv=True, **z["cam"])
    scale = np.abs(z["deriv"]).max()
    assert np.abs(bw["deriv"] - z["deriv"]).max() <= 1e-11 * scale
    gp = np.concatenate([bw["g_position"], bw["g_orientation"], [bw["g_inv_scale"]]])
    assert np.abs(gp - z["g_pose"]).max() <= 1e-11 * np.abs(z["g_pose"]).max()
    assert np.abs(bw["g_sdf"] - z["g_sdf_exact"]).max() <= 1e-11 * np.abs(z["g_sdf_exact"]).max()


@pytest.mark.parametrize("name", NAMES)
def test_f32_oracle_tracks_reference(name):
    """The float32 instantiation (CUDA-kernel arithmetic) stays within fp32 rounding of the
    float64 reference on the golden scenes: same hit mask, same step counts."""
    z = load_golden(name)
    d, steps, _ = oracle.render(z["sdf"], z["position"], z["orientation"], z["inv_scale"],
                                z["W"], z["H"], threshold=float(z["threshold"]),
                                dtype=np.float32, extras=True, **z["cam"])
    gold = z["depth"]
    hit = gold > 0
    assert ((d > 0) == hit).all()
    assert (np.abs(d - gold)[hit] / gold[hit]).max() < 1e-5
    assert (np.where(hit, steps, 0) == z["steps"]).all()
    bw = oracle.render_backward(z["g"], d, z["sdf"], z["position"], z["orientation"],
                                z["inv_scale"], z["W"], z["H"], sdf_grad_mode="exact",
                                dtype=np.float32, **z["cam"])
    gp = np.concatenate([bw["g_position"], bw["g_orientation"], [bw["g_inv_scale"]]])
    assert np.abs(gp - z["g_pose"]).max() <= 1e-4 * np.abs(z["g_pose"]).max()
    assert np.abs(bw["g_sdf"] - z["g_sdf_exact"]).max() <= 1e-4 * np.abs(z["g_sdf_exact"]).max()


def test_reference_weight_list_differs_from_exact():
    """the CUDA kernel's corner-weight list is a permutation, not a rounding effect."""
    z = load_golden("mug_z0_r64")
    kw = dict(dtype=np.float64, **z["cam"])
    args = (z["g"], z["depth"], z["sdf"], z["position"], z["orientation"], z["inv_scale"],
            z["W"], z["H"])
    ref = oracle.render_backward(*args, sdf_grad_mode="reference", **kw)["g_sdf"]
    exact = oracle.render_backward(*args, sdf_grad_mode="exact", **kw)["g_sdf"]
    rel = np.linalg.norm(ref - exact) / np.linalg.norm(exact)
    assert rel > 0.3
    # pose gradients do not depend on the mode
    a = oracle.render_backward(*args, sdf_grad_mode="reference", **kw)
    b = oracle.render_backward(*args, sdf_grad_mode="exact", **kw)
    assert np.array_equal(a["g_position"], b["g_position"])


def test_threads_do_not_change_the_result():
    z = load_golden("torus_r32")
    a = oracle.render(z["sdf"], z["position"], z["orientation"], z["inv_scale"], z["W"], z["H"],
                      threshold=0.003, nthreads=1, **z["cam"])
    b = oracle.render(z["sdf"], z["position"], z["orientation"], z["inv_scale"], z["W"], z["H"],
                      threshold=0.003, nthreads=4, **z["cam"])
    assert np.array_equal(a, b)


def test_composite_and_l1_helpers():
    layers = np.array([[[0.0, 2.0], [3.0, 0.0]], [[1.0, 1.5], [3.0, 0.0]]])
    depth, winner = oracle.composite_min_depth(layers)
    assert np.array_equal(depth, [[1.0, 1.5], [3.0, 0.0]])
    assert np.array_equal(winner, [[1, 1], [0, -1]])
    est = np.array([[1.0, 0.0], [2.0, 3.0]])
    obs = np.array([[1.5, 1.0], [0.0, 3.0]])
    loss, grad, n = oracle.l1_depth_loss(est, obs)
    assert n == 2 and loss == pytest.approx(0.25)
    assert np.array_equal(grad, [[-0.5, 0.0], [0.0, 0.0]])
    assert oracle.l1_depth_loss(np.zeros((2, 2)), obs)[2] == 0


def test_result_selection_oracle_matches_the_reference_methods():
    """tests/golden/selection.npz was produced by the reference's own _compute_inlier_ratio /
    _update_best_estimate (simple_setup.py:177-211, source executed unchanged by
    tests/golden/make_golden_selection.py): the numpy oracle and the package's torch statement
    reproduce its ratios and its running best exactly; the reference's RETURNED estimate is the last
    iterate (it keeps references to the live tensors), ours the copy of the best iteration."""
    import os

    import torch

    from oracle import hypothesis_step as hs
    from sdfest_b200.differentiable_renderer import Camera
    from util import GOLDEN_DIR
    from sdfest_b200.estimation import HypothesisOptimizer

    z = np.load(os.path.join(GOLDEN_DIR, "selection.npz"))
    obs, est, thr = z["obs"], z["est"], float(z["threshold"])
    n, (H, W) = est.shape[0], obs.shape
    best = hs.BestEstimate()
    cam = Camera(W, H, 30.0, 30.0, W / 2, H / 2, pixel_center=0.5)
    opt = HypothesisOptimizer(cam, 0.005, torch.tensor(obs), torch.zeros(1, 3), torch.tensor([[0.0, 0, 0, 1]]),
                              torch.ones(1), sdf=torch.zeros(1, 8, 8, 8), inlier_threshold=thr)
    for it in range(n):
        ni, nv = hs.inlier_counts(obs, est[it], thr)
        r = best.update(ni, nv, it + 1, (z["positions"][it],))
        assert np.float32(r) == z["ratios"][it]
        assert np.float32(best.ratio) == z["best_so_far"][it]
        with torch.no_grad():
            opt.position.copy_(torch.tensor(z["positions"][it])[None])
        opt._track_best_torch(torch.tensor(est[it])[None])
        assert float(opt.inlier_ratio[0]) == float(z["ratios"][it])
        assert float(opt.best_inlier_ratio[0]) == float(z["best_so_far"][it])
    assert best.iteration == 4 and int(opt.best_iteration[0]) == 4
    np.testing.assert_array_equal(best.params[0], z["positions"][3])
    np.testing.assert_array_equal(opt.result("best_inlier_ratio")[0][0].numpy(), z["positions"][3])
    # what the reference hands back for "best_inlier_ratio" is the last iterate
    np.testing.assert_array_equal(z["returned_position"], z["last_position"])
    np.testing.assert_allclose(opt.result("last_iteration")[0][0].numpy(), z["last_position"], rtol=1e-6)
