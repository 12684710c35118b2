"""The exact per-ray functions the sm_100a kernels inline (sdfest_b200/csrc/sdfr_core.cuh),
compiled for the host with g++ and checked against the oracle -- no GPU needed.  Covers the
arithmetic and the conservativeness of the projected-box rectangle culling; the warp/CTA glue
(shuffles, atomics, barriers) is covered by the -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from util import (default_camera, golden_names, load_golden, mug_sdf, sdf_box, sdf_sphere,
                  sdf_torus, shoemake)

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cpp")
SO = os.path.join(HERE, "host_emul", "libemul.so")
SO_REF = os.path.join(HERE, "host_emul", "libemul_ref.so")
CORE = os.path.join(os.path.dirname(HERE), "sdfest_b200", "csrc", "sdfr_core.cuh")
f32 = np.float32


def _build(so, *defines):
    if (not os.path.isfile(so)
            or os.path.getmtime(so) < max(os.path.getmtime(SRC), os.path.getmtime(CORE))):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-ffp-contract=off", *defines, "-fPIC", "-shared", SRC, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def emul():
    """The per-ray functions in the REFERENCE'S operation order (-DSDFR_REFERENCE_ROUNDING):
    bit-identical to the fp32 oracle."""
    return _build(SO_REF, "-DSDFR_REFERENCE_ROUNDING")


@pytest.fixture(scope="module")
def emul_product():
    """The per-ray functions as the shipped library compiles them: the march evaluates the cell
    coordinate with one fma per axis and fma lerps (sdfr_core.cuh march<>), a few ulp away from the
    reference's expression."""
    return _build(SO)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_forward(lib, sdf, pos, quat, inv_scale, W, H, cam, thr, use_rect=1):
    sdf = np.ascontiguousarray(sdf, f32)
    pos, quat = np.asarray(pos, f32), np.asarray(quat, f32)
    s = np.asarray([inv_scale], f32)
    depth = np.empty((H, W), f32)
    steps = np.empty((H, W), np.int32)
    rect = np.zeros(8, np.int32)
    cf = [ctypes.c_float(v) for v in (cam["cx"], cam["cy"], cam["fx"], cam["fy"], thr)]
    lib.emul_forward(P(sdf), sdf.shape[0], P(pos), P(quat), P(s), W, H, *cf, P(depth), P(steps),
                     P(rect), use_rect)
    return depth, steps, rect


@pytest.mark.parametrize("name", golden_names())
def test_device_math_matches_oracle_on_golden(emul, name):
    z = load_golden(name)
    thr = float(z["threshold"])
    depth, steps, _ = run_forward(emul, z["sdf"], z["position"], z["orientation"],
                                  float(z["inv_scale"]), z["W"], z["H"], z["cam"], thr)
    d_or, st_or, _ = oracle.render(z["sdf"], z["position"], z["orientation"], z["inv_scale"],
                                   z["W"], z["H"], threshold=thr, extras=True, **z["cam"])
    # same operation order, no FMA contraction on the host -> bit-identical
    assert np.array_equal(depth, d_or)
    assert np.array_equal(steps, st_or)
    # and within fp32 rounding of the reference's float64 renderer
    hit = z["depth"] > 0
    assert ((depth > 0) == hit).all()
    assert (np.abs(depth - z["depth"])[hit] / z["depth"][hit]).max() < 1e-5


@pytest.mark.parametrize("name", golden_names())
def test_product_march_is_within_rounding_of_the_oracle(emul_product, name):
    """Default build (fma march): same hit mask; depth within 2e-6 relative of the fp32 oracle for
    every pixel whose step count agrees (a termination test decided within an ulp may flip and move
    that pixel by up to `threshold` relative: at most 0.1 % of the hit pixels); within 1e-5 of the
    reference's float64 renderer like the reference-order build."""
    z = load_golden(name)
    thr = float(z["threshold"])
    depth, steps, _ = run_forward(emul_product, z["sdf"], z["position"], z["orientation"],
                                  float(z["inv_scale"]), z["W"], z["H"], z["cam"], thr)
    d_or, st_or, _ = oracle.render(z["sdf"], z["position"], z["orientation"], z["inv_scale"],
                                   z["W"], z["H"], threshold=thr, extras=True, **z["cam"])
    hit = d_or > 0
    assert ((depth > 0) == hit).mean() > 0.999
    both = hit & (depth > 0)
    rel = np.abs(depth - d_or)[both] / d_or[both]
    same = (steps == st_or)[both]
    assert same.mean() > 0.999
    assert rel[same].max() < 2e-6
    assert rel.max() < 1.5 * thr
    ref_hit = z["depth"] > 0
    ok = ref_hit & (depth > 0)
    rel64 = np.abs(depth - z["depth"])[ok] / z["depth"][ok]
    assert (rel64 < 1e-5).mean() > 0.999 and rel64.max() < 1.5 * thr


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_device_backward_matches_oracle(emul, name, mode):
    z = load_golden(name)
    sdf = np.ascontiguousarray(z["sdf"], f32)
    R = sdf.shape[0]
    pos, quat = z["position"].astype(f32), z["orientation"].astype(f32)
    s = np.asarray([z["inv_scale"]], f32)
    depth = oracle.render(sdf, pos, quat, s, z["W"], z["H"], threshold=float(z["threshold"]),
                          **z["cam"])
    g = z["g"].astype(f32)
    bw = oracle.render_backward(g, depth, sdf, pos, quat, s, z["W"], z["H"], sdf_grad_mode=mode,
                                **z["cam"])
    gs, gp = np.zeros((R, R, R)), np.zeros(8)
    cf = [ctypes.c_float(z["cam"][k]) for k in ("cx", "cy", "fx", "fy")]
    emul.emul_backward(P(g), P(depth), P(sdf), R, P(pos), P(quat), P(s), z["W"], z["H"], *cf,
                       int(mode == "exact"), P(gs), P(gp))
    gpo = np.concatenate([bw["g_position"], bw["g_orientation"], [bw["g_inv_scale"]]])
    assert np.abs(gp - gpo).max() <= 1e-6 * np.abs(gpo).max()
    assert np.abs(gs - bw["g_sdf"]).max() <= 1e-5 * np.abs(bw["g_sdf"]).max()
    if mode == "exact":  # and against the reference's own derivatives
        assert np.abs(gp - z["g_pose"]).max() <= 1e-4 * np.abs(z["g_pose"]).max()
        assert np.abs(gs - z["g_sdf_exact"]).max() <= 1e-4 * np.abs(z["g_sdf_exact"]).max()


def test_rectangle_culling_is_conservative(emul):
    """Random poses incl. objects partly off-screen, partly behind the camera, camera inside the
    box: rendering with and without the projected-box rectangle must give identical images."""
    rng = np.random.default_rng(0)
    grids = [sdf_sphere(16), sdf_box(20), sdf_torus(24)]
    W, H = 96, 64
    cam = default_camera(W, H)
    n_culled = 0
    for i in range(60):
        sdf = grids[i % 3]
        scale = rng.uniform(0.05, 0.6)
        pos = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.6, 0.6), rng.uniform(-1.5, 0.3)])
        q = shoemake(100 + i)
        a, _, rect = run_forward(emul, sdf, pos, q, 1 / scale, W, H, cam, 0.005, use_rect=1)
        b, _, _ = run_forward(emul, sdf, pos, q, 1 / scale, W, H, cam, 0.005, use_rect=0)
        assert np.array_equal(a, b), (i, pos, scale, rect)
        n_culled += (rect[2] - rect[0]) * (rect[3] - rect[1]) < W * H
    assert n_culled > 10  # the culling is actually exercised


def test_hull_culling_is_conservative(emul):
    """Same as above for the silhouette (convex hull of the projected corners) culling that the
    kernels apply per 8x4-pixel warp tile: identical images, and it culls more than the rectangle."""
    rng = np.random.default_rng(1)
    grids = [sdf_sphere(16), sdf_box(20), sdf_torus(24)]
    W, H = 96, 64
    cam = default_camera(W, H)
    culled, with_hull = 0, 0
    for i in range(120):
        sdf = grids[i % 3]
        scale = rng.uniform(0.05, 0.6)
        pos = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.6, 0.6), rng.uniform(-1.5, 0.3)])
        q = shoemake(500 + i) if i % 4 else np.array([0.0, 0.0, 0.0, 1.0])  # incl. face-on views
        a, _, rect = run_forward(emul, sdf, pos, q, 1 / scale, W, H, cam, 0.005, use_rect=2)
        b, _, _ = run_forward(emul, sdf, pos, q, 1 / scale, W, H, cam, 0.005, use_rect=0)
        assert np.array_equal(a, b), (i, pos, scale, rect)
        assert 0 <= rect[4] <= 8
        with_hull += rect[4] >= 3
        culled += int(rect[5])
    assert with_hull > 40 and culled > 5000


def test_full_size_mug_frame_matches_oracle(emul):
    """One 640x480 frame of the reference workload (mug SDF, default camera)."""
    sdf = mug_sdf()
    W, H = 640, 480
    cam = default_camera(W, H)
    pos, q, scale = [0.02, -0.01, -0.4], shoemake(1), 0.15
    depth, steps, rect = run_forward(emul, sdf, pos, q, 1 / scale, W, H, cam, 0.005, use_rect=2)
    assert rect[4] >= 4 and rect[5] > 0.2 * (rect[2] - rect[0]) * (rect[3] - rect[1])
    d_or, st_or, _ = oracle.render(sdf, pos, q, 1 / scale, W, H, threshold=0.005, extras=True,
                                   nthreads=8, **cam)
    assert np.array_equal(depth, d_or)
    assert (depth > 0).sum() > 20000
    assert (rect[2] - rect[0]) * (rect[3] - rect[1]) < W * H


# ------------------------------------------------------------------------------------------
# round 2: pose gradients through moment sums; empty-space bounds
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_moment_backward_matches_oracle(emul, name, mode):
    """The kernels accumulate 13 moment sums and map them to the 8 pose gradients once
    (sdfr_core.cuh: pixel_backward_moments / moments_to_pose): same mathematics as cu:391-457, a
    different order of fp32 operations -- 1e-4 of the largest component against the oracle (the
    shipped tolerance is 1e-3) and against the reference's own float64 derivatives."""
    z = load_golden(name)
    sdf = np.ascontiguousarray(z["sdf"], f32)
    R = sdf.shape[0]
    pos, quat = z["position"].astype(f32), z["orientation"].astype(f32)
    s = np.asarray([z["inv_scale"]], f32)
    depth = oracle.render(sdf, pos, quat, s, z["W"], z["H"], threshold=float(z["threshold"]), **z["cam"])
    g = z["g"].astype(f32)
    bw = oracle.render_backward(g, depth, sdf, pos, quat, s, z["W"], z["H"], sdf_grad_mode=mode, **z["cam"])
    gs, gp = np.zeros((R, R, R)), np.zeros(8)
    cf = [ctypes.c_float(z["cam"][k]) for k in ("cx", "cy", "fx", "fy")]
    emul.emul_backward_moments(P(g), P(depth), P(sdf), R, P(pos), P(quat), P(s), z["W"], z["H"], *cf,
                               int(mode == "exact"), P(gs), P(gp))
    gpo = np.concatenate([bw["g_position"], bw["g_orientation"], [bw["g_inv_scale"]]])
    for sl in (slice(0, 3), slice(3, 7), slice(7, 8)):  # per parameter group, as the GPU tests gate them
        assert np.abs(gp[sl] - gpo[sl]).max() <= 1e-4 * np.abs(gpo[sl]).max(), (sl, gp[sl], gpo[sl])
    assert np.abs(gs - bw["g_sdf"]).max() <= 1e-5 * np.abs(bw["g_sdf"]).max()
    if mode == "exact":
        ref = z["g_pose"]
        for sl in (slice(0, 3), slice(3, 7), slice(7, 8)):
            assert np.abs(gp[sl] - ref[sl]).max() <= 2e-4 * np.abs(ref[sl]).max()


def _bounds_record(sdf, pos, inv_scale, thr, lib):
    from oracle import grid_bounds as gb

    tau = gb.hit_tau(pos, inv_scale, thr)
    lib.emul_hit_tau.restype = ctypes.c_float
    dev_tau = lib.emul_hit_tau(P(np.asarray(pos, f32)), ctypes.c_float(inv_scale), ctypes.c_float(thr))
    assert np.float32(dev_tau) == tau  # the oracle's bound is the device function's, bit for bit
    lo, hi = gb.cell_bounds(sdf, tau)
    return gb.pack(lo, hi, tau), lo, hi


def run_forward_bounds(lib, sdf, pos, quat, inv_scale, W, H, cam, thr, rec, use_rect=2):
    sdf = np.ascontiguousarray(sdf, f32)
    pos, quat = np.asarray(pos, f32), np.asarray(quat, f32)
    s = np.asarray([inv_scale], f32)
    depth = np.empty((H, W), f32)
    steps = np.empty((H, W), np.int32)
    rect = np.zeros(8, np.int32)
    cf = [ctypes.c_float(v) for v in (cam["cx"], cam["cy"], cam["fx"], cam["fy"], thr)]
    lib.emul_forward_bounds(P(sdf), sdf.shape[0], P(pos), P(quat), P(s), W, H, *cf, P(depth), P(steps),
                            P(rect), use_rect, None if rec is None else P(rec))
    return depth, steps, rect


def test_empty_space_bounds_never_change_the_image(emul):
    """Random poses / shapes / thresholds (objects partly off-screen, behind the camera, camera inside
    the box, thresholds up to 0.05): rendering with the cell bounds of oracle/grid_bounds.py is
    bit-identical -- depth AND step counts of every traced ray -- to rendering without, and it
    marches far fewer rays."""
    from util import sdf_bottle, sdf_bowl

    rng = np.random.default_rng(7)
    grids = [sdf_sphere(16), sdf_box(20), sdf_torus(24), sdf_bottle(32), sdf_bowl(28), mug_sdf()]
    W, H = 96, 64
    cam = default_camera(W, H)
    marched_with, marched_without = 0, 0
    for i in range(90):
        sdf = grids[i % len(grids)]
        scale = rng.uniform(0.05, 0.6)
        pos = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.6, 0.6), rng.uniform(-1.5, 0.3)], f32)
        q = shoemake(900 + i) if i % 5 else np.array([0.0, 0.0, 0.0, 1.0])
        thr = [0.0, 0.003, 0.005, 0.01, 0.05][i % 5]
        rec, lo, hi = _bounds_record(sdf, pos, 1 / scale, thr, emul)
        a, sa, ra = run_forward_bounds(emul, sdf, pos, q, 1 / scale, W, H, cam, thr, rec)
        b, sb, rb = run_forward_bounds(emul, sdf, pos, q, 1 / scale, W, H, cam, thr, None)
        assert np.array_equal(a, b), (i, pos, scale, thr, lo, hi)
        assert np.array_equal(sa[a > 0], sb[a > 0])
        marched_with += int(ra[6])
        marched_without += int(rb[6])
    assert marched_without > 0 and marched_with < 0.7 * marched_without, (marched_with, marched_without)


def test_empty_space_bounds_full_size_mug(emul_product):
    """The reference workload (640x480, mug, default camera), shipped arithmetic: identical image,
    and the bounds prove most of the box's rays empty."""
    sdf = mug_sdf()
    W, H = 640, 480
    cam = default_camera(W, H)
    pos, q, scale, thr = np.array([0.02, -0.01, -0.4], f32), shoemake(1), 0.15, 0.005
    rec, lo, hi = _bounds_record(sdf, pos, 1 / scale, thr, emul_product)
    a, sa, ra = run_forward_bounds(emul_product, sdf, pos, q, 1 / scale, W, H, cam, thr, rec)
    b, sb, rb = run_forward_bounds(emul_product, sdf, pos, q, 1 / scale, W, H, cam, thr, None)
    assert np.array_equal(a, b)
    assert (a > 0).sum() > 20000
    assert ra[6] < 0.75 * rb[6], (ra[6], rb[6], lo, hi)
    # bounds computed for a SMALLER threshold bound are ignored (frame_pose checks tau), never trusted
    from oracle import grid_bounds as gb

    weak = gb.pack(lo, hi, np.float32(1e-6))
    c, _, rc = run_forward_bounds(emul_product, sdf, pos, q, 1 / scale, W, H, cam, thr, weak)
    assert np.array_equal(c, b) and rc[6] == rb[6]


def test_empty_grid_bounds_zero_the_image(emul):
    """A field that never comes near zero: lo > hi, the rectangle is empty, nothing is marched."""
    from oracle import grid_bounds as gb

    sdf = np.full((16, 16, 16), 0.5, f32)
    pos, thr = np.array([0.0, 0.0, -1.0], f32), 0.005
    tau = gb.hit_tau(pos, 1 / 0.3, thr)
    lo, hi = gb.cell_bounds(sdf, tau)
    assert (lo > hi).all()
    a, _, ra = run_forward_bounds(emul, sdf, pos, np.array([0, 0, 0, 1.0]), 1 / 0.3, 96, 64,
                                  default_camera(96, 64), thr, gb.pack(lo, hi, tau))
    assert not a.any() and ra[6] == 0 and ra[2] - ra[0] == 0
