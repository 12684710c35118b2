"""Parity AT THE BENCHMARKED CONFIGURATIONS, through the C ABI, against the f32 oracle and the
reference's own CUDA extension (oracle/_ref).

 * C2 -- exactly what ``bench.py`` times: ``sdfr_skew_grids`` -> ``sdfr_compare_fused`` (kernel
   instantiation ``sdfr_forward_kernel<64, skewed, MODE 2>``, one grid per hypothesis, 64 hypotheses x
   640x480) -> ``sdfr_scale_grads``.  Depth, ``loss_sum``, ``n_overlap``, ``n_inlier`` and all four
   gradients of EVERY hypothesis are compared with the oracle (sdf_renderer_cuda.cu:241-468 restated)
   and with the reference extension's forward + torch masked L1 (simple_setup.py:125-131) + backward.
 * C3 -- 16 objects x 128^3 composited into one 1280x720 depth map against ``oracle.render_composite``.
 * C4 -- the sweep's shape mix: hypotheses of one category share ONE grid (``sdf_stride = 0``), pose
   gradients only.

Tolerances are BASELINE.json's: depth 1e-5 relative (>= 99.9 % of the hit pixels; the rest are
one-step termination flips bounded by the threshold), gradients 1e-3 relative.  The observed slack is
written to ``gpurun_out/parity_stats.json`` (quoted in DESIGN.md section 2).
"""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import build_ref
from sdfest_b200 import _lib
from sdfest_b200 import synthetic as syn
from sdfest_b200.differentiable_renderer import Camera, render_depth_batched, render_depth_composite
from util import cell_face_mask, depth_parity, grad_close, pose_grad_parity, sdf_grad_parity

pytestmark = pytest.mark.gpu

W, H, FX, FY, CX, CY = 640, 480, 320.0, 320.0, 320.0, 240.0  # bench.py / default.yaml:1-8
R, THR = 64, 0.005
CAM = dict(cx=CX, cy=CY, fx=FX, fy=FY)
DEPTH_RTOL, GRAD_RTOL = 1e-5, 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STATS = {}


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    return v


def _record(key, value):
    _STATS[key] = _plain(value)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_stats.json"), "w") as f:
            json.dump(_STATS, f, indent=1)
    except OSError:
        pass


def _merge(infos):
    """Worst case / totals over a list of depth_parity dictionaries."""
    out = dict(n=0, good=0, mask_flips=0, step_flips=0, max_rel_good=0.0)
    for i in infos:
        for k in ("n", "good", "mask_flips", "step_flips"):
            out[k] += i[k]
        out["max_rel_good"] = max(out["max_rel_good"], i["max_rel_good"])
    return out


@pytest.fixture(scope="module")
def ref_ext():
    mod = build_ref.load_module()
    if mod is None:
        pytest.skip("oracle/_ref/sdf_renderer_cpp.so not present")
    return mod


@pytest.fixture(scope="module")
def c2(cuda_device):
    """The bench's C2 step (bench.py::main: skew -> fused -> scale) run once, results on the host."""
    dev = cuda_device
    B = 64
    lib = _lib.lib()
    cam = Camera(W, H, FX, FY, CX, CY, pixel_center=0.5)
    hyp = syn.make_hypotheses(B, seed=0, device=dev)
    grids = syn.hypothesis_grids(hyp["shape_param"], R, dev)
    pos, quat, inv_s = hyp["position"], hyp["orientation"], hyp["inv_scale"]
    base = syn.make_hypotheses(1, seed=0, device=dev)
    obs = render_depth_batched(syn.hypothesis_grids(base["shape_param"], R, dev), base["position"],
                               base["orientation"], base["inv_scale"], THR, cam)[0].contiguous()
    stream = torch.cuda.current_stream().cuda_stream
    RRR = R ** 3
    n_sk = ctypes.c_longlong(0)
    _lib.check(lib.sdfr_skewed_pitches(R, None, None, ctypes.byref(n_sk)), "sdfr_skewed_pitches")
    SK = int(n_sk.value)
    skewed = torch.full((B, SK), float("nan"), device=dev)  # the padding must never be read
    depth = torch.full((B, H, W), float("nan"), device=dev)
    sums = torch.full((3, B), float("nan"), device=dev)  # loss_sum, n_overlap, n_inlier
    g_sdf = torch.full((B, R, R, R), float("nan"), device=dev)
    g_pos, g_quat = torch.full((B, 3), float("nan"), device=dev), torch.full((B, 4), float("nan"), device=dev)
    g_is = torch.full((B,), float("nan"), device=dev)
    flags = _lib.GRAD_ALL | _lib.ZERO_GRADS
    _lib.check(lib.sdfr_skew_grids(grids.data_ptr(), R, RRR, B, skewed.data_ptr(), SK, stream), "skew")
    _lib.check(lib.sdfr_compare_fused_inliers(
        skewed.data_ptr(), R, SK, _lib.LAYOUT_SKEWED, pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(),
        B, W, H, CX, CY, FX, FY, THR, obs.data_ptr(), 0, depth.data_ptr(), sums[0].data_ptr(),
        sums[1].data_ptr(), 0.03, sums[2].data_ptr(), g_sdf.data_ptr(), RRR, g_pos.data_ptr(),
        g_quat.data_ptr(), g_is.data_ptr(), flags, None, stream), "sdfr_compare_fused_inliers")
    raw = [t.clone() for t in (g_sdf, g_pos, g_quat, g_is)]
    _lib.check(lib.sdfr_scale_grads(sums[1].data_ptr(), None, R, B, g_sdf.data_ptr(), RRR, g_pos.data_ptr(),
                                    g_quat.data_ptr(), g_is.data_ptr(), _lib.GRAD_ALL, None, 0, stream), "scale")
    torch.cuda.synchronize()
    return dict(B=B, dev=dev, grids=grids, pos=pos, quat=quat, inv_s=inv_s, obs=obs, depth=depth, sums=sums,
                g_sdf=g_sdf, g_pos=g_pos, g_quat=g_quat, g_is=g_is, raw=raw, skewed=skewed, SK=SK,
                np=dict(grids=grids.cpu().numpy(), pos=pos.cpu().numpy(), quat=quat.cpu().numpy(),
                        inv_s=inv_s.cpu().numpy(), obs=obs.cpu().numpy(), depth=depth.cpu().numpy(),
                        sums=sums.cpu().numpy(), g_sdf=g_sdf.cpu().numpy(), g_pos=g_pos.cpu().numpy(),
                        g_quat=g_quat.cpu().numpy(), g_is=g_is.cpu().numpy()))


def test_c2_plain_fused_entry_point_equals_the_inlier_variant(c2):
    """bench.py calls sdfr_compare_fused; the fixture called sdfr_compare_fused_inliers (the same
    kernel instantiation with one more counter): identical depth and sums, gradients to atomics order."""
    lib, dev, B = _lib.lib(), c2["dev"], c2["B"]
    depth = torch.empty_like(c2["depth"])
    sums = torch.empty(2, B, device=dev)
    g = [torch.empty_like(t) for t in c2["raw"]]
    _lib.check(lib.sdfr_compare_fused(
        c2["skewed"].data_ptr(), R, c2["SK"], _lib.LAYOUT_SKEWED, c2["pos"].data_ptr(), c2["quat"].data_ptr(),
        c2["inv_s"].data_ptr(), B, W, H, CX, CY, FX, FY, THR, c2["obs"].data_ptr(), 0, depth.data_ptr(),
        sums[0].data_ptr(), sums[1].data_ptr(), g[0].data_ptr(), R ** 3, g[1].data_ptr(), g[2].data_ptr(),
        g[3].data_ptr(), _lib.GRAD_ALL | _lib.ZERO_GRADS, None, torch.cuda.current_stream().cuda_stream),
        "sdfr_compare_fused")
    torch.cuda.synchronize()
    assert torch.equal(depth, c2["depth"])
    assert torch.equal(sums[1], c2["sums"][1])
    grad_close(sums[0].cpu().numpy(), c2["sums"][0].cpu().numpy(), 1e-6, "loss_sum")
    for a, b, nm in zip(g, c2["raw"], ("sdf", "position", "orientation", "inv_scale")):
        grad_close(a.cpu().numpy(), b.cpu().numpy(), 1e-4, "raw " + nm)


def test_c2_benchmarked_path_matches_oracle(c2):
    n = c2["np"]
    infos, errs = [], dict(position=0.0, sdf=0.0, loss=0.0)  # "position": worst pose group, relative to max|ref|
    for b in range(c2["B"]):
        d_gpu = n["depth"][b]
        d_or = oracle.render(n["grids"][b], n["pos"][b], n["quat"][b], n["inv_s"][b], W, H, threshold=THR,
                             nthreads=os.cpu_count() or 1, **CAM)
        infos.append(depth_parity(d_gpu, d_or, THR, rtol=DEPTH_RTOL))
        # loss / counters / gradients from the GPU's own depth, so that a one-step termination flip
        # (already gated above) cannot leak into these gates
        loss, g, n_over = oracle.l1_depth_loss(d_gpu, n["obs"])
        assert n_over == int(n["sums"][1, b]), (b, n_over, n["sums"][1, b])
        assert n_over > 1000
        assert abs(n["sums"][0, b] - loss * n_over) <= 1e-5 * loss * n_over
        errs["loss"] = max(errs["loss"], abs(n["sums"][0, b] - loss * n_over) / max(loss * n_over, 1e-30))
        obs32, d32 = n["obs"].astype(np.float32), d_gpu.astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.abs(obs32 - d32) / obs32  # simple_setup.py:183, IEEE float32 like the kernel
        inl = int(((rel < np.float32(0.03)) & (obs32 > 0) & (d32 > 0)).sum())
        assert inl == int(n["sums"][2, b]), (b, inl, n["sums"][2, b])
        bw = oracle.render_backward(g.astype(np.float32), d_gpu, n["grids"][b], n["pos"][b], n["quat"][b],
                                    n["inv_s"][b], W, H, sdf_grad_mode="reference", want_deriv=True,
                                    nthreads=os.cpu_count() or 1, **CAM)
        errs["position"] = max(errs["position"], pose_grad_parity(
            (n["g_pos"][b], n["g_quat"][b], n["g_is"][b]), bw, g, GRAD_RTOL, f"hypothesis {b}",
            position=n["pos"][b], inv_scale=n["inv_s"][b]) * GRAD_RTOL)
        e, n_bad = sdf_grad_parity(n["g_sdf"][b], bw["g_sdf"], GRAD_RTOL, np.abs(g).max() / n["inv_s"][b], what=f"sdf[{b}]")
        errs["sdf"] = max(errs["sdf"], e)
        errs["sdf_cell_flip_voxels"] = errs.get("sdf_cell_flip_voxels", 0) + n_bad
    _record("c2_vs_oracle", dict(depth=_merge(infos), grad_max_rel_err=errs))


def test_c2_benchmarked_path_matches_reference_cuda_extension(c2, ref_ext):
    """The loop bench.py::time_reference_extension times, as a checker: per hypothesis the reference's
    forward, the pipeline's masked L1 in torch, the reference's backward."""
    infos, errs = [], dict(position=0.0, orientation=0.0, inv_scale=0.0, sdf=0.0)
    obs = c2["obs"]
    for b in range(c2["B"]):
        sdf, p, q, s = c2["grids"][b], c2["pos"][b], c2["quat"][b], c2["inv_s"][b:b + 1]
        (d_ref,) = ref_ext.forward(sdf, p, q, s, W, H, CX, CY, FX, FY, THR)
        ours = c2["depth"][b]
        infos.append(depth_parity(ours.cpu().numpy(), d_ref.cpu().numpy(), THR, rtol=DEPTH_RTOL))
        mask = (obs > 0) & (ours > 0)
        g = torch.where(mask, torch.sign(ours - obs), torch.zeros_like(ours)) / mask.sum()
        r_sdf, r_p, r_q, r_s = ref_ext.backward(g.contiguous(), ours.contiguous(), sdf, p, q, s, W, H, CX, CY,
                                                FX, FY)
        errs["position"] = max(errs["position"], grad_close(c2["g_pos"][b].cpu().numpy(), r_p.cpu().numpy(),
                                                            GRAD_RTOL, f"position[{b}] vs ref ext"))
        errs["orientation"] = max(errs["orientation"],
                                  grad_close(c2["g_quat"][b].cpu().numpy(), r_q.cpu().numpy(), GRAD_RTOL,
                                             f"orientation[{b}] vs ref ext"))
        errs["inv_scale"] = max(errs["inv_scale"],
                                grad_close(c2["g_is"][b:b + 1].cpu().numpy(), r_s.cpu().numpy(), GRAD_RTOL,
                                           f"inv_scale[{b}] vs ref ext"))
        e, n_bad = sdf_grad_parity(c2["g_sdf"][b].cpu().numpy(), r_sdf.cpu().numpy(), GRAD_RTOL,
                                   float(g.abs().max() / c2["inv_s"][b]), what=f"sdf[{b}] vs ref ext")
        errs["sdf"] = max(errs["sdf"], e)
        errs["sdf_cell_flip_voxels"] = errs.get("sdf_cell_flip_voxels", 0) + n_bad
    _record("c2_vs_reference_ext", dict(depth=_merge(infos), grad_max_rel_err=errs))


# ------------------------------------------------------------------------------------------
# C3: 16 objects x 128^3, one 1280x720 frame
# ------------------------------------------------------------------------------------------
def c3_scene(dev):
    """scripts/gpu_configs.py::config3 / bench.py --config c3."""
    K, R3 = 16, 128
    names = ("mug", "bowl", "bottle")
    grids = torch.stack([syn.category_grid(names[k % 3], R3, "cpu", shape_param=0.3 * ((k % 5) - 2) / 2)
                         for k in range(K)]).contiguous().to(dev)
    g = torch.Generator().manual_seed(3)
    z = -(0.6 + 0.6 * torch.rand(K, generator=g))
    ix, iy = torch.arange(K) % 4, torch.arange(K) // 4
    pos = torch.stack([(ix - 1.5) * 0.42 * (-z), (iy - 1.5) * 0.24 * (-z), z], 1)
    quat = syn.random_unit_quaternions(K, g)
    scale = 0.08 + 0.07 * torch.rand(K, generator=g)
    pos, quat, inv_s = (t.float().contiguous().to(dev) for t in (pos, quat, 1.0 / scale))
    return grids, pos, quat, inv_s


def test_c3_composite_at_size_matches_oracle(cuda_device):
    K, W3, H3 = 16, 1280, 720
    cam_d = dict(cx=640.0, cy=360.0, fx=640.0, fy=640.0)
    cam = Camera(W3, H3, 640.0, 640.0, 640.0, 360.0, pixel_center=0.5)
    grids, pos, quat, inv_s = c3_scene(cuda_device)
    a = [t.clone().requires_grad_(True) for t in (grids, pos, quat, inv_s)]
    depth, winner = render_depth_composite(*a, THR, cam)
    ng, npos, nq, ns = (t.cpu().numpy() for t in (grids, pos, quat, inv_s))
    nt = os.cpu_count() or 1
    d_or, w_or = oracle.render_composite(ng, npos, nq, ns, W3, H3, threshold=THR, nthreads=nt, **cam_d)
    info = depth_parity(depth.detach().cpu().numpy(), d_or, THR, rtol=DEPTH_RTOL)
    w_gpu = winner.cpu().numpy()
    assert (w_gpu == w_or).mean() > 0.9995
    assert len(set(np.unique(w_gpu)) - {-1}) >= 12, "most of the 16 objects must be visible"
    g = np.random.default_rng(4).standard_normal((H3, W3)).astype(np.float32)
    # no upstream gradient where the hit point sits on a cell face (tests/util.py::cell_face_mask)
    d_first = depth.detach().cpu().numpy()
    n_masked = 0
    for k in range(K):
        face = cell_face_mask(np.where(w_gpu == k, d_first, 0), npos[k], nq[k], ns[k], 128, **cam_d)
        g[face] = 0.0
        n_masked += int(face.sum())
    assert n_masked < 0.05 * (w_gpu >= 0).sum()
    depth.backward(torch.as_tensor(g, device=cuda_device))
    d_np = depth.detach().cpu().numpy()
    errs = dict(position=0.0, sdf=0.0)
    for k in range(K):
        if not (w_gpu == k).any():
            continue
        dk = np.where(w_gpu == k, d_np, 0).astype(np.float32)  # oracle.render_composite_backward, with derivatives
        bw = oracle.render_backward(g, dk, ng[k], npos[k], nq[k], ns[k], W3, H3, sdf_grad_mode="reference",
                                    want_deriv=True, nthreads=nt, **cam_d)
        bws = {k: bw}
        gk = np.where(dk != 0, g, 0)
        errs["position"] = max(errs["position"], pose_grad_parity(
            (a[1].grad[k].cpu().numpy(), a[2].grad[k].cpu().numpy(), a[3].grad[k].cpu().numpy()), bw, gk,
            GRAD_RTOL, f"obj{k}", position=npos[k], inv_scale=ns[k]) * GRAD_RTOL)
        e, n_bad = sdf_grad_parity(a[0].grad[k].cpu().numpy(), bws[k]["g_sdf"], GRAD_RTOL,
                                   float(np.abs(g).max() / ns[k]), what=f"obj{k} sdf")
        errs["sdf"] = max(errs["sdf"], e)
        errs["sdf_cell_flip_voxels"] = errs.get("sdf_cell_flip_voxels", 0) + n_bad
    _record("c3_vs_oracle", dict(depth=info, grad_max_rel_err=errs,
                                 winner_agreement=float((w_gpu == w_or).mean())))


# ------------------------------------------------------------------------------------------
# C4: the sweep's shape mix -- one shared grid per category, pose gradients only
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("category", syn.CATEGORIES)
def test_c4_shared_grid_pose_sweep_matches_oracle(cuda_device, category):
    from sdfest_b200.differentiable_renderer.sdf_renderer import _grid_operand

    dev, B, lib = cuda_device, 24, _lib.lib()
    cam = Camera(W, H, FX, FY, CX, CY, pixel_center=0.5)
    hyp = syn.make_hypotheses(3 * B, seed=0, device=dev)  # scripts/gpu_sweep.py: hypothesis i has category i % 3
    sel = torch.arange(3 * B, device=dev)[syn.CATEGORIES.index(category)::3]
    pos, quat, inv_s = (hyp[k][sel].contiguous() for k in ("position", "orientation", "inv_scale"))
    grid = syn.category_grid(category, R, dev)[None].contiguous()
    base = syn.make_hypotheses(1, seed=0, device=dev)
    obs = render_depth_batched(syn.category_grid("mug", R, dev)[None], base["position"], base["orientation"],
                               base["inv_scale"], THR, cam)[0].contiguous()
    src, stride, layout = _grid_operand(grid, R, 0, B, W * H)
    assert stride == 0 and layout == _lib.LAYOUT_SKEWED
    depth = torch.empty(B, H, W, device=dev)
    sums = torch.empty(2, B, device=dev)
    g_pos, g_quat, g_is = torch.empty(B, 3, device=dev), torch.empty(B, 4, device=dev), torch.empty(B, device=dev)
    flags = _lib.GRAD_POSITION | _lib.GRAD_ORIENTATION | _lib.GRAD_INV_SCALE
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.sdfr_compare_fused(
        src.data_ptr(), R, 0, layout, pos.data_ptr(), quat.data_ptr(), inv_s.data_ptr(), B, W, H, CX, CY, FX, FY,
        THR, obs.data_ptr(), 0, depth.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), None, 0,
        g_pos.data_ptr(), g_quat.data_ptr(), g_is.data_ptr(), flags | _lib.ZERO_GRADS, None, st), "sdfr_compare_fused")
    _lib.check(lib.sdfr_scale_grads(sums[1].data_ptr(), None, R, B, None, 0, g_pos.data_ptr(), g_quat.data_ptr(),
                                    g_is.data_ptr(), flags, None, 0, st), "sdfr_scale_grads")
    torch.cuda.synchronize()
    ngrid, nobs = grid[0].cpu().numpy(), obs.cpu().numpy()
    infos = []
    for b in range(B):
        p, q, s = pos[b].cpu().numpy(), quat[b].cpu().numpy(), float(inv_s[b])
        d_gpu = depth[b].cpu().numpy()
        d_or = oracle.render(ngrid, p, q, s, W, H, threshold=THR, nthreads=os.cpu_count() or 1, **CAM)
        infos.append(depth_parity(d_gpu, d_or, THR, rtol=DEPTH_RTOL))
        loss, g, n_over = oracle.l1_depth_loss(d_gpu, nobs)
        assert n_over == int(sums[1, b])
        if n_over == 0:
            assert float(g_pos[b].abs().max()) == 0.0
            continue
        assert abs(float(sums[0, b]) - loss * n_over) <= 1e-5 * loss * n_over
        bw = oracle.render_backward(g.astype(np.float32), d_gpu, ngrid, p, q, s, W, H, want_sdf=False,
                                    want_deriv=True, nthreads=os.cpu_count() or 1, **CAM)
        pose_grad_parity((g_pos[b].cpu().numpy(), g_quat[b].cpu().numpy(), float(g_is[b])), bw, g, GRAD_RTOL,
                         f"{category}[{b}]", position=p, inv_scale=s)
    _record(f"c4_{category}_vs_oracle", dict(depth=_merge(infos)))


# ------------------------------------------------------------------------------------------
# empty-space bounds (sdfr_grid_bounds): exact cell bounds, and renders that do not change
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["dense", "skewed"])
def test_grid_bounds_kernel_matches_oracle(c2, layout):
    from oracle import grid_bounds as gb

    lib, B, dev = _lib.lib(), c2["B"], c2["dev"]
    st = torch.cuda.current_stream().cuda_stream
    src, stride, lay = (c2["grids"], R ** 3, _lib.LAYOUT_DENSE) if layout == "dense" else \
        (c2["skewed"], c2["SK"], _lib.LAYOUT_SKEWED)
    out = torch.full((B, 8), -7, dtype=torch.int32, device=dev)
    _lib.check(lib.sdfr_grid_bounds(src.data_ptr(), R, stride, lay, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(),
                                    B, THR, out.data_ptr(), st), "sdfr_grid_bounds")
    got = out.cpu().numpy()
    n = c2["np"]
    for b in range(B):
        tau = gb.hit_tau(n["pos"][b], n["inv_s"][b], THR)
        lo, hi = gb.cell_bounds(n["grids"][b], tau)
        assert got[b, 6:7].view(np.float32)[0] == tau
        assert (got[b, 0:3] == lo).all() and (got[b, 3:6] == hi).all(), (b, got[b], lo, hi)
    # one grid shared by the batch: tau is the largest of the batch
    shared = torch.full((1, 8), -7, dtype=torch.int32, device=dev)
    _lib.check(lib.sdfr_grid_bounds(src.data_ptr(), R, 0, lay, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(),
                                    B, THR, shared.data_ptr(), st), "sdfr_grid_bounds")
    got = shared.cpu().numpy()[0]
    tau = max(gb.hit_tau(n["pos"][b], n["inv_s"][b], THR) for b in range(B))
    lo, hi = gb.cell_bounds(n["grids"][0], tau)
    assert got[6:7].view(np.float32)[0] == tau and (got[0:3] == lo).all() and (got[3:6] == hi).all()


@pytest.mark.parametrize("layout", ["dense", "skewed"])
def test_bounds_from_slab_minima_equal_the_scan(c2, layout):
    """Fixed grids: sdfr_grid_slab_minima once + sdfr_bounds_from_minima per pose == sdfr_grid_bounds, for
    the C2 grids at two thresholds, a shared grid, and an odd resolution with negative and positive
    minima; the minima themselves against numpy."""
    lib, B, dev = _lib.lib(), c2["B"], c2["dev"]
    st = torch.cuda.current_stream().cuda_stream
    src, stride, lay = (c2["grids"], R ** 3, _lib.LAYOUT_DENSE) if layout == "dense" else \
        (c2["skewed"], c2["SK"], _lib.LAYOUT_SKEWED)
    minima = torch.full((B, 3, R), float("nan"), device=dev)
    _lib.check(lib.sdfr_grid_slab_minima(src.data_ptr(), R, stride, lay, B, minima.data_ptr(), st), "minima")
    g = c2["grids"].view(B, R, R, R)
    want = torch.stack([g.amin(dim=(2, 3)), g.amin(dim=(1, 3)), g.amin(dim=(1, 2))], 1)
    assert torch.equal(minima, want)
    for thr in (THR, 0.05):
        a = torch.full((B, 8), -7, dtype=torch.int32, device=dev)
        b = torch.full((B, 8), -9, dtype=torch.int32, device=dev)
        _lib.check(lib.sdfr_grid_bounds(src.data_ptr(), R, stride, lay, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(),
                                        B, thr, a.data_ptr(), st), "scan")
        _lib.check(lib.sdfr_bounds_from_minima(minima.data_ptr(), R, B, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(),
                                               B, thr, b.data_ptr(), st), "from minima")
        assert torch.equal(a[:, :7], b[:, :7])
    a1 = torch.full((1, 8), -7, dtype=torch.int32, device=dev)
    b1 = torch.full((1, 8), -9, dtype=torch.int32, device=dev)
    _lib.check(lib.sdfr_grid_bounds(src.data_ptr(), R, 0, lay, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(), B, THR,
                                    a1.data_ptr(), st), "scan shared")
    _lib.check(lib.sdfr_bounds_from_minima(minima.data_ptr(), R, 1, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(), B,
                                           THR, b1.data_ptr(), st), "from minima shared")
    assert torch.equal(a1[:, :7], b1[:, :7])
    if layout == "dense":
        R2 = 37
        g2 = (torch.rand(2, R2, R2, R2, device=dev) - 0.5) * torch.tensor([1.0, 1e-3], device=dev).view(2, 1, 1, 1) + \
            torch.tensor([0.45, 0.0], device=dev).view(2, 1, 1, 1)
        g2 = g2.contiguous()
        m2 = torch.empty((2, 3, R2), device=dev)
        _lib.check(lib.sdfr_grid_slab_minima(g2.data_ptr(), R2, R2 ** 3, _lib.LAYOUT_DENSE, 2, m2.data_ptr(), st), "minima")
        assert torch.equal(m2, torch.stack([g2.amin(dim=(2, 3)), g2.amin(dim=(1, 3)), g2.amin(dim=(1, 2))], 1))
        a2 = torch.empty((2, 8), dtype=torch.int32, device=dev)
        b2 = torch.empty((2, 8), dtype=torch.int32, device=dev)
        _lib.check(lib.sdfr_grid_bounds(g2.data_ptr(), R2, R2 ** 3, 0, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(), 2,
                                        THR, a2.data_ptr(), st), "scan")
        _lib.check(lib.sdfr_bounds_from_minima(m2.data_ptr(), R2, 2, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(), 2,
                                               THR, b2.data_ptr(), st), "from minima")
        assert torch.equal(a2[:, :7], b2[:, :7])
    assert lib.sdfr_grid_slab_minima(None, R, 0, 0, 1, None, None) == -1
    assert lib.sdfr_grid_slab_minima(None, R, 0, 0, 0, None, None) == 0
    assert lib.sdfr_bounds_from_minima(minima.data_ptr(), R, 3, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(), B, THR,
                                       b1.data_ptr(), None) == -2


def test_skew_and_bounds_in_one_pass(c2):
    """sdfr_skew_grids_bounds == sdfr_skew_grids + sdfr_grid_bounds (what bench.py's step launches)."""
    lib, B, dev = _lib.lib(), c2["B"], c2["dev"]
    st = torch.cuda.current_stream().cuda_stream
    RRR, SK = R ** 3, c2["SK"]
    sep_sk = torch.full((B, SK), -7.0, device=dev)
    sep_b = torch.empty((B, 8), dtype=torch.int32, device=dev)
    _lib.check(lib.sdfr_skew_grids(c2["grids"].data_ptr(), R, RRR, B, sep_sk.data_ptr(), SK, st), "skew")
    _lib.check(lib.sdfr_grid_bounds(sep_sk.data_ptr(), R, SK, _lib.LAYOUT_SKEWED, c2["pos"].data_ptr(),
                                    c2["inv_s"].data_ptr(), B, THR, sep_b.data_ptr(), st), "bounds")
    one_sk = torch.full((B, SK), -7.0, device=dev)
    one_b = torch.full((B, 8), -3, dtype=torch.int32, device=dev)
    _lib.check(lib.sdfr_skew_grids_bounds(c2["grids"].data_ptr(), R, RRR, B, one_sk.data_ptr(), SK,
                                          c2["pos"].data_ptr(), c2["inv_s"].data_ptr(), THR, one_b.data_ptr(), st),
               "sdfr_skew_grids_bounds")
    torch.cuda.synchronize()
    assert torch.equal(one_sk, sep_sk)
    assert torch.equal(one_b[:, :7], sep_b[:, :7])
    # odd resolution, shared grid
    R2 = 37
    g = torch.rand(R2, R2, R2, device=dev) - 0.3
    n_sk = ctypes.c_longlong(0)
    _lib.check(lib.sdfr_skewed_pitches(R2, None, None, ctypes.byref(n_sk)), "pitches")
    a_sk, b_sk = torch.full((1, n_sk.value), -7.0, device=dev), torch.full((1, n_sk.value), -7.0, device=dev)
    a_b, b_b = torch.empty((1, 8), dtype=torch.int32, device=dev), torch.empty((1, 8), dtype=torch.int32, device=dev)
    _lib.check(lib.sdfr_skew_grids(g.data_ptr(), R2, 0, 1, a_sk.data_ptr(), n_sk.value, st), "skew")
    _lib.check(lib.sdfr_grid_bounds(g.data_ptr(), R2, 0, _lib.LAYOUT_DENSE, c2["pos"].data_ptr(), c2["inv_s"].data_ptr(),
                                    B, 0.0, a_b.data_ptr(), st), "bounds")
    _lib.check(lib.sdfr_skew_grids_bounds(g.data_ptr(), R2, 0, B, b_sk.data_ptr(), n_sk.value, c2["pos"].data_ptr(),
                                          c2["inv_s"].data_ptr(), 0.0, b_b.data_ptr(), st), "skew+bounds")
    torch.cuda.synchronize()
    assert torch.equal(a_sk, b_sk) and torch.equal(a_b[:, :7], b_b[:, :7])
    from oracle import grid_bounds as gb

    lo, hi = gb.cell_bounds(g.cpu().numpy(), float(a_b[0, 6:7].view(torch.float32)))
    assert (a_b[0, 0:3].cpu().numpy() == lo).all() and (a_b[0, 3:6].cpu().numpy() == hi).all()


def test_c2_with_empty_space_bounds_is_unchanged(c2):
    """bench.py's step with the bounds pass in front: depth, counters bit-identical; sums and raw
    gradients identical up to the order of the fp32 atomics; and far fewer rays are marched."""
    from sdfest_b200.differentiable_renderer import forward_stats

    lib, dev, B = _lib.lib(), c2["dev"], c2["B"]
    st = torch.cuda.current_stream().cuda_stream
    bounds = torch.empty((B, 8), dtype=torch.int32, device=dev)
    _lib.check(lib.sdfr_grid_bounds(c2["skewed"].data_ptr(), R, c2["SK"], _lib.LAYOUT_SKEWED, c2["pos"].data_ptr(),
                                    c2["inv_s"].data_ptr(), B, THR, bounds.data_ptr(), st), "sdfr_grid_bounds")
    depth = torch.full_like(c2["depth"], float("nan"))
    sums = torch.empty(3, B, device=dev)
    g = [torch.empty_like(t) for t in c2["raw"]]
    _lib.check(lib.sdfr_compare_fused_inliers(
        c2["skewed"].data_ptr(), R, c2["SK"], _lib.LAYOUT_SKEWED, c2["pos"].data_ptr(), c2["quat"].data_ptr(),
        c2["inv_s"].data_ptr(), B, W, H, CX, CY, FX, FY, THR, c2["obs"].data_ptr(), 0, depth.data_ptr(),
        sums[0].data_ptr(), sums[1].data_ptr(), 0.03, sums[2].data_ptr(), g[0].data_ptr(), R ** 3,
        g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), _lib.GRAD_ALL | _lib.ZERO_GRADS, bounds.data_ptr(), st),
        "sdfr_compare_fused_inliers")
    torch.cuda.synchronize()
    assert torch.equal(depth, c2["depth"])
    assert torch.equal(sums[1:], c2["sums"][1:])
    grad_close(sums[0].cpu().numpy(), c2["sums"][0].cpu().numpy(), 1e-6, "loss_sum")
    for a, b, nm in zip(g, c2["raw"], ("sdf", "position", "orientation", "inv_scale")):
        grad_close(a.cpu().numpy(), b.cpu().numpy(), 1e-4, "raw " + nm + " with bounds")
    # unfused forward + backward with bounds == without
    d2 = torch.full_like(depth, float("nan"))
    _lib.check(lib.sdfr_forward(c2["grids"].data_ptr(), R, R ** 3, _lib.LAYOUT_DENSE, c2["pos"].data_ptr(),
                                c2["quat"].data_ptr(), c2["inv_s"].data_ptr(), B, W, H, CX, CY, FX, FY, THR,
                                d2.data_ptr(), bounds.data_ptr(), st), "sdfr_forward")
    assert torch.equal(d2, c2["depth"])
    cam = Camera(W, H, FX, FY, CX, CY, pixel_center=0.5)
    full = forward_stats(c2["grids"], c2["pos"], c2["quat"], c2["inv_s"], THR, cam)
    tight = forward_stats(c2["grids"], c2["pos"], c2["quat"], c2["inv_s"], THR, cam, empty_space=True)
    assert tight["hit_pixels"] == full["hit_pixels"]
    assert tight["box_pixels"] < 0.5 * full["box_pixels"] and tight["samples"] < 0.8 * full["samples"]
    _record("c2_empty_space", dict(full=full, with_bounds=tight))


def test_scale_grads_inside_the_bounds_box_equals_the_full_pass(c2):
    """bench.py's step: skew + bounds -> fused render with those bounds -> sdfr_scale_grads over the bounds box
    only.  Same gradients as scaling the whole grid (every voxel outside the box is an exact zero), and the
    box really is a small part of the grid."""
    lib, dev, B = _lib.lib(), c2["dev"], c2["B"]
    st = torch.cuda.current_stream().cuda_stream
    bounds = torch.empty((B, 8), dtype=torch.int32, device=dev)
    skewed = torch.empty_like(c2["skewed"])
    _lib.check(lib.sdfr_skew_grids_bounds(c2["grids"].data_ptr(), R, R ** 3, B, skewed.data_ptr(), c2["SK"],
                                          c2["pos"].data_ptr(), c2["inv_s"].data_ptr(), THR, bounds.data_ptr(), st), "skew")
    out = []
    for use_box in (False, True):
        depth = torch.empty_like(c2["depth"])
        sums = torch.empty(2, B, device=dev)
        g = [torch.full_like(t, float("nan")) for t in c2["raw"]]
        _lib.check(lib.sdfr_compare_fused(
            skewed.data_ptr(), R, c2["SK"], _lib.LAYOUT_SKEWED, c2["pos"].data_ptr(), c2["quat"].data_ptr(),
            c2["inv_s"].data_ptr(), B, W, H, CX, CY, FX, FY, THR, c2["obs"].data_ptr(), 0, depth.data_ptr(),
            sums[0].data_ptr(), sums[1].data_ptr(), g[0].data_ptr(), R ** 3, g[1].data_ptr(), g[2].data_ptr(),
            g[3].data_ptr(), _lib.GRAD_ALL | _lib.ZERO_GRADS, bounds.data_ptr(), st), "fused")
        _lib.check(lib.sdfr_scale_grads(sums[1].data_ptr(), None, R, B, g[0].data_ptr(), R ** 3, g[1].data_ptr(),
                                        g[2].data_ptr(), g[3].data_ptr(), _lib.GRAD_ALL,
                                        bounds.data_ptr() if use_box else None, 1 if use_box else 0, st), "scale")
        torch.cuda.synchronize()
        out.append(g)
    for a, b, nm in zip(out[0], out[1], ("sdf", "position", "orientation", "inv_scale")):
        assert bool(torch.isfinite(b).all())
        grad_close(b.cpu().numpy(), a.cpu().numpy(), 1e-4, "box-scaled " + nm)
    bb = bounds.cpu().numpy()
    ext = np.prod(bb[:, 3:6] - bb[:, 0:3] + 2, axis=1) / R ** 3
    assert 0.02 < float(ext.mean()) < 0.5
    # nothing outside the box was non-zero to begin with
    g_sdf = out[1][0].view(B, R, R, R)
    for b in range(0, B, 9):
        lo, hi = bb[b, 0:3], bb[b, 3:6]
        inside = torch.zeros(R, R, R, dtype=torch.bool, device=dev)
        inside[lo[0]:hi[0] + 2, lo[1]:hi[1] + 2, lo[2]:hi[2] + 2] = True
        assert float(g_sdf[b][~inside].abs().max()) == 0.0
    # one shared grid: bounds_stride 0 is accepted, other strides are not
    assert lib.sdfr_scale_grads(sums[1].data_ptr(), None, R, B, None, 0, out[1][1].data_ptr(), None, None,
                                _lib.GRAD_POSITION, bounds.data_ptr(), 0, st) == 0
    assert lib.sdfr_scale_grads(sums[1].data_ptr(), None, R, B, None, 0, out[1][1].data_ptr(), None, None,
                                _lib.GRAD_POSITION, bounds.data_ptr(), 2, st) == -2


def test_c3_composite_with_bounds_is_unchanged(cuda_device):
    from sdfest_b200.differentiable_renderer import set_empty_space_policy

    cam = Camera(1280, 720, 640.0, 640.0, 640.0, 360.0, pixel_center=0.5)
    grids, pos, quat, inv_s = c3_scene(cuda_device)
    try:
        set_empty_space_policy("off")
        d0, w0 = render_depth_composite(grids, pos, quat, inv_s, THR, cam)
        set_empty_space_policy("on")
        d1, w1 = render_depth_composite(grids, pos, quat, inv_s, THR, cam)
    finally:
        set_empty_space_policy("auto")
    assert torch.equal(d0, d1) and torch.equal(w0, w1)
