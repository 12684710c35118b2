"""Host-side logic of the batched render-and-compare loop, on CPU: the batched pc_loss against
a golden vector from the reference's own estimation/losses.py, observed-point lifting, and the
hypothesis sharding / loss all-gather over a 2-rank gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdfest_b200.differentiable_renderer import Camera
from oracle.pc_loss import pc_loss, point_loss  # torch restatement of the reference's loss (test infrastructure)
from sdfest_b200.estimation import (depth_to_pointcloud, depth_to_pointclouds, gather_losses, global_best,
                                    shard_range, subsample_points)
from util import GOLDEN_DIR


def test_pc_loss_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "pcloss_torus16.npz"))
    t = lambda k, g=False: torch.tensor(z[k], dtype=torch.float64, requires_grad=g)
    pos, quat, scale, sdf = t("position", True), t("orientation", True), t("scale", True), t("sdf", True)
    val = pc_loss(t("points"), pos[None], quat[None], scale[None], sdf[None])
    assert val.shape == (1, 200)
    assert np.abs(val[0].detach().numpy() - z["value"]).max() < 1e-12
    assert (val == 0).sum() == (z["value"] == 0).sum() > 0
    val[0].abs().mean().backward()
    for got, key in ((pos.grad, "g_position"), (quat.grad, "g_orientation"), (scale.grad, "g_scale"),
                     (sdf.grad, "g_sdf")):
        assert np.abs(got.numpy() - z[key]).max() <= 1e-10 * max(np.abs(z[key]).max(), 1e-30)


def test_pc_loss_batches_are_independent():
    z = np.load(os.path.join(GOLDEN_DIR, "pcloss_torus16.npz"))
    t = lambda k: torch.tensor(z[k], dtype=torch.float32)
    B = 3
    pos = t("position")[None] + 0.01 * torch.arange(B)[:, None]
    quat = t("orientation")[None].repeat(B, 1)
    scale = t("scale").reshape(1).repeat(B) * torch.tensor([1.0, 1.1, 0.9])
    shared = pc_loss(t("points"), pos, quat, scale, t("sdf")[None])
    for b in range(B):
        one = pc_loss(t("points"), pos[b:b + 1], quat[b:b + 1], scale[b:b + 1], t("sdf")[None])
        assert torch.allclose(shared[b], one[0], atol=1e-6)


def test_depth_to_pointcloud_opengl_convention():
    cam = Camera(4, 3, 2.0, 2.0, 2.0, 1.5, pixel_center=0.5)  # -> cx=1.5, cy=1.0 at centre 0
    depth = torch.zeros(3, 4)
    depth[1, 3] = 2.0
    depth[0, 0] = 1.0
    pts = depth_to_pointcloud(depth, cam)
    assert pts.shape == (2, 3)
    # row-major order of nonzero(): (0,0) first
    assert torch.allclose(pts[0], torch.tensor([(0 - 1.5) * 1.0 / 2, -(0 - 1.0) * 1.0 / 2, -1.0]))
    assert torch.allclose(pts[1], torch.tensor([(3 - 1.5) * 2.0 / 2, -(1 - 1.0) * 2.0 / 2, -2.0]))


def test_padded_instance_clouds_keep_every_instance_mean():
    """K instances with clouds of different sizes: rows are the instance's own cloud followed by padding
    that pc_loss maps to exactly 0, so sum / count is the reference's mean over the instance's points."""
    z = np.load(os.path.join(GOLDEN_DIR, "pcloss_torus16.npz"))
    t = lambda k: torch.tensor(z[k], dtype=torch.float32)  # noqa: E731
    cam = Camera(16, 12, 14.0, 14.0, 8.0, 6.0, pixel_center=0.5)
    g = torch.Generator().manual_seed(3)
    depth = torch.zeros(3, 12, 16)
    depth[0, 2:9, 3:12] = 0.9 + 0.05 * torch.rand(7, 9, generator=g)
    depth[1, 5:7, 5:8] = 1.0  # instance 2 has no valid pixel at all
    clouds, counts = depth_to_pointclouds(depth, cam)
    assert counts.tolist() == [63, 6, 0] and clouds.shape == (3, 63, 3)
    for k in range(3):
        own = depth_to_pointcloud(depth[k], cam)
        assert torch.equal(clouds[k, : counts[k]], own)
        assert bool((clouds[k, counts[k]:] > 1e5).all())
    capped, n_capped = depth_to_pointclouds(depth, cam, max_points=10)
    assert n_capped.tolist() == [10, 6, 0] and capped.shape == (3, 10, 3)
    assert torch.equal(capped[0], subsample_points(depth_to_pointcloud(depth[0], cam), 10))
    # the padding contributes exactly 0 to the point loss of any pose
    pos = torch.tensor([[0.0, 0.0, -0.95]]).repeat(3, 1)
    quat = t("orientation")[None].repeat(3, 1)
    scale = torch.full((3,), 0.4)
    padded = point_loss(clouds, pos, quat, scale, t("sdf")[None])  # sum / M over the padded rows
    for k in range(2):
        own = point_loss(clouds[k, : counts[k]], pos[k:k + 1], quat[k:k + 1], scale[k:k + 1], t("sdf")[None])
        assert float(own) > 0
        assert torch.allclose(padded[k] * clouds.shape[1] / counts[k], own[0], rtol=1e-6)
    assert float(padded[2]) == 0.0
    with pytest.raises(RuntimeError):
        depth_to_pointclouds(depth[0], cam)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        losses = torch.arange(lo, hi, dtype=torch.float32) * 0.5 + 1.0
        if rank == 1:
            losses[0] = float("nan")  # a hypothesis without overlap must never win
            losses[-1] = 0.25         # the global best lives on rank 1
        full = gather_losses(losses)
        # shard sizes known up front (shard_range): same result, no size exchange per call
        known = [h - l for l, h in (shard_range(n_total, r, world) for r in range(world))]
        again = gather_losses(losses, sizes=known)
        assert torch.equal(torch.nan_to_num(full), torch.nan_to_num(again))
        try:
            gather_losses(losses, sizes=[1] * world if known != [1] * world else [2] * world)
            raise AssertionError("wrong shard sizes must be rejected")
        except ValueError:
            pass
        idx, val = global_best(losses, lo)
        q.put((rank, full.tolist(), idx, val))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_two_rank_gloo_gather_and_argmin(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [i * 0.5 + 1.0 for i in range(n_total)]
    lo1 = shard_range(n_total, 1, 2)[0]
    expect[-1] = 0.25
    for rank, full, idx, val in out:
        assert len(full) == n_total
        for i, (a, b) in enumerate(zip(full, expect)):
            assert (np.isnan(a) and i == lo1) or a == b
        assert idx == n_total - 1 and val == 0.25


def test_hypothesis_optimizer_instance_bookkeeping_on_cpu():
    """Host side of object-instance batching (no kernels run): hypothesis b gets the depth map, the
    cloud and the point count of instance[b]; malformed instance arguments are rejected."""
    from sdfest_b200.estimation import HypothesisOptimizer

    cam = Camera(16, 12, 14.0, 14.0, 8.0, 6.0, pixel_center=0.5)
    depth = torch.zeros(2, 12, 16)
    depth[0, 2:6, 3:9] = 1.0
    depth[1, 5:7, 5:8] = 1.2
    B = 5
    instance = torch.tensor([0, 0, 1, 1, 0])
    pos, quat, scale = torch.zeros(B, 3), torch.tensor([[0.0, 0, 0, 1]]).repeat(B, 1), torch.ones(B)
    sdf = torch.zeros(1, 8, 8, 8)
    opt = HypothesisOptimizer(cam, 0.005, depth, pos, quat, scale, sdf=sdf, instance=instance)
    assert opt.optimizer_impl == "torch"
    assert opt.point_counts.tolist() == [24, 24, 6, 6, 24]
    assert opt.points.shape == (B, 24, 3) and opt.depth_obs.shape == (B, 12, 16)
    assert torch.equal(opt.depth_obs[2], depth[1]) and torch.equal(opt.depth_obs[4], depth[0])
    assert torch.equal(opt.points[3, :6], depth_to_pointcloud(depth[1], cam))
    # default: one observation per hypothesis
    one_each = HypothesisOptimizer(cam, 0.005, depth, pos[:2], quat[:2], scale[:2], sdf=sdf)
    assert one_each.point_counts.tolist() == [24, 6]
    # a single shared observation keeps the (M,3) cloud and has no per-hypothesis counts
    shared = HypothesisOptimizer(cam, 0.005, depth[0], pos, quat, scale, sdf=sdf)
    assert shared.point_counts is None and shared.points.shape == (24, 3)
    for bad in (dict(depth_obs=depth[0], instance=instance),            # instance without (K,H,W)
                dict(depth_obs=depth, instance=instance[:3]),             # wrong length
                dict(depth_obs=depth, instance=instance + 1),             # out of range
                dict(depth_obs=depth, instance=-instance)):
        with pytest.raises(ValueError):
            HypothesisOptimizer(cam, 0.005, bad["depth_obs"], pos, quat, scale, sdf=sdf, instance=bad["instance"])


def _stand_in_render_and_compare(sdf, position, orientation, inv_scale, depth_obs, threshold, camera):
    B = position.shape[0]
    target = depth_obs[depth_obs > 0].mean()
    loss = (position[:, 2] + target) ** 2 + 0.1 * (orientation[:, 3] - 1) ** 2 + 0.01 * inv_scale
    return loss, depth_obs[None].expand(B, -1, -1) * 1.01, torch.ones(B)


def _loop_inputs(n_total):
    cam = Camera(16, 12, 14.0, 14.0, 8.0, 6.0, pixel_center=0.5)
    obs = torch.zeros(12, 16)
    obs[3:8, 4:10] = 0.9
    g = torch.Generator().manual_seed(4)
    pos = torch.tensor([[0.0, 0.0, -1.0]]) + 0.05 * torch.randn(n_total, 3, generator=g)
    quat = torch.nn.functional.normalize(torch.tensor([[0.0, 0, 0, 1]]) + 0.1 * torch.randn(n_total, 4, generator=g), dim=1)
    scale = 0.3 + 0.02 * torch.rand(n_total, generator=g)
    sdf = torch.rand(1, 8, 8, 8, generator=g) - 0.3
    return cam, obs, pos, quat, scale, sdf


def _loop_worker(rank, world, port, n_total, q):
    from sdfest_b200.estimation import HypothesisOptimizer, hypotheses, losses

    hypotheses.render_and_compare = _stand_in_render_and_compare  # the CUDA renderer's stand-in on CPU
    losses.point_loss = point_loss  # and the CUDA point loss's (the product has no CPU path)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cam, obs, pos, quat, scale, sdf = _loop_inputs(n_total)
        lo, hi = shard_range(n_total, rank, world)
        opt = HypothesisOptimizer(cam, 0.005, obs, pos[lo:hi], quat[lo:hi], scale[lo:hi], sdf=sdf,
                                  inlier_threshold=0.03)
        gathered = opt.run(4, gather_every=2)
        q.put((rank, gathered.tolist(), opt.position.detach().tolist()))
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_sharded_loop_equals_the_single_process_loop():
    """HypothesisOptimizer.run over two gloo ranks with an uneven shard (3 + 2 hypotheses): every rank
    ends with all five losses, in hypothesis order, equal to the unsharded run (hypotheses are
    independent: the only exchange is the loss all-gather)."""
    n_total = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 77
    procs = [ctx.Process(target=_loop_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single_q = ctx.Queue()
    one = ctx.Process(target=_loop_worker, args=(0, 1, port + 1, n_total, single_q))
    one.start()
    _, want, want_pos = single_q.get(timeout=180)
    one.join(timeout=60)
    assert len(want) == n_total
    for rank, losses_all, _ in out:
        np.testing.assert_allclose(losses_all, want, rtol=1e-6)
    np.testing.assert_allclose(out[0][2] + out[1][2], want_pos, rtol=1e-6)


def test_runtime_overview_has_the_reference_yaml_layout():
    """runtime_analysis.reference_overview: the document of the reference's generate_runtime_overview
    (estimation/scripts/real_data.py:286-319): same section and phase names, seconds, per-run statistics."""
    from sdfest_b200.estimation import runtime_analysis as ra

    w = {"decode": 0.5, "render_compare": 0.2, "point_loss": 0.1, "backward": 0.6, "optimizer": 0.05, "total": 1.45}
    wo = {"decode": 0.0, "render_compare": 0.2, "point_loss": 0.1, "backward": 0.1, "optimizer": 0.05, "total": 0.45}
    doc = ra.reference_overview(w, wo, iterations_per_run=50, runs=4, init_ms=6.0, config={"dataset": "synthetic"})
    assert set(doc) == {"dataset", "results_with_decode", "results_without_decode"}
    a, b = doc["results_with_decode"], doc["results_without_decode"]
    assert set(a) == {"init", "decode", "render", "losses", "backward", "optimizer"}
    assert "decode" not in b and set(b) == set(a) - {"decode"}
    for sec in (a, b):
        for st in sec.values():
            assert set(st) == {"total", "total_calls", "mean", "calls_per_run", "total_per_run"}
            assert abs(st["total"] - st["mean"] * st["total_calls"]) < 1e-12
    assert a["render"]["mean"] == 0.2e-3 and a["render"]["total_calls"] == 200 and a["render"]["calls_per_run"] == 50.0
    assert abs(a["backward"]["total_per_run"] - 0.6e-3 * 50) < 1e-12
    assert a["init"] == {"total": 0.024, "total_calls": 4, "mean": 0.006, "calls_per_run": 1.0, "total_per_run": 0.006}
