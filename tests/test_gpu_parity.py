"""GPU parity tests: the sm_100a kernels (through the C ABI / ctypes binding) against
 (1) the golden vectors generated from the reference's CPU renderer,
 (2) the CPU oracle on the same seeded inputs,
 (3) the reference's own CUDA extension (oracle/_ref, compiled from /root/reference) on the same
     GPU, when the prebuilt binary is present.
Tolerances are BASELINE.json's: depth 1e-5 relative, gradients 1e-3 relative."""
import numpy as np
import pytest
import torch

import oracle
from oracle import build_ref
from sdfest_b200 import _lib
from sdfest_b200.differentiable_renderer import (Camera, forward_stats, render_and_compare,
                                                 render_depth_batched, render_depth_composite,
                                                 render_depth_gpu)
from util import (default_camera, depth_parity, golden_names, grad_close, load_golden, mug_sdf,
                  sdf_bottle, sdf_bowl, sdf_box, sdf_sphere, sdf_torus, shoemake)

pytestmark = pytest.mark.gpu

DEPTH_RTOL = 1e-5  # BASELINE.json north_star: depth to 1e-5 relative
GRAD_RTOL = 1e-3   # BASELINE.json north_star: gradients within 1e-3 relative


def cam_obj(W, H, cam):
    return Camera(W, H, cam["fx"], cam["fy"], cam["cx"], cam["cy"], pixel_center=0.5)


def T(a, dev, grad=False):
    a = np.asarray(a, dtype=np.float32)
    t = torch.as_tensor(np.ascontiguousarray(a).reshape(a.shape), device=dev)  # keeps 0-dim
    return t.requires_grad_(grad)


@pytest.fixture(scope="module")
def ref_ext():
    mod = build_ref.load_module()
    if mod is None:
        pytest.skip("oracle/_ref/sdf_renderer_cpp.so not present")
    return mod


def test_native_library_is_loaded(cuda_device):
    lib = _lib.lib()
    assert b"sm_100a" in lib.sdfr_build_info()
    maps = open("/proc/self/maps").read()
    assert "libsdfrender.so" in maps


# ------------------------------------------------------------------------------------------
# (1) golden vectors from the reference's CPU renderer
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_forward_matches_reference_golden(cuda_device, name):
    z = load_golden(name)
    depth = render_depth_gpu(T(z["sdf"], cuda_device), T(z["position"], cuda_device),
                             T(z["orientation"], cuda_device), T([z["inv_scale"]], cuda_device),
                             z["W"], z["H"], float(z["fov_deg"]), float(z["threshold"]))
    assert depth.shape == (z["H"], z["W"]) and depth.dtype == torch.float32
    depth_parity(depth.cpu().numpy(), z["depth"], float(z["threshold"]), rtol=DEPTH_RTOL)


@pytest.mark.parametrize("name", golden_names())
def test_backward_matches_reference_golden_exact_mode(cuda_device, name):
    z = load_golden(name)
    sdf = T(z["sdf"], cuda_device, True)
    p = T(z["position"], cuda_device, True)
    q = T(z["orientation"], cuda_device, True)
    s = T([z["inv_scale"]], cuda_device, True)
    depth = render_depth_gpu(sdf, p, q, s, z["W"], z["H"], float(z["fov_deg"]),
                             float(z["threshold"]), sdf_grad_mode="exact")
    # only meaningful where the hit masks agree (they do on the golden scenes)
    assert ((depth.detach().cpu().numpy() > 0) == (z["depth"] > 0)).all()
    depth.backward(T(z["g"], cuda_device))
    gp = np.concatenate([p.grad.cpu().numpy(), q.grad.cpu().numpy(), s.grad.cpu().numpy()])
    grad_close(gp, z["g_pose"], GRAD_RTOL, "pose grads vs reference derivatives")
    grad_close(sdf.grad.cpu().numpy(), z["g_sdf_exact"], GRAD_RTOL, "sdf grads vs reference")


# ------------------------------------------------------------------------------------------
# (2) CPU oracle, larger seeded scenes
# ------------------------------------------------------------------------------------------
SCENES = {
    # name: (sdf builder, R, position, quat seed, scale, W, H, threshold)
    "mug_default_view": (mug_sdf, 64, [0.02, -0.01, -0.4], 1, 0.15, 640, 480, 0.005),
    "mug_small_far": (mug_sdf, 64, [0.1, 0.05, -0.9], 3, 0.055, 640, 480, 0.003),
    "torus_r32_odd_image": (lambda: sdf_torus(32), 32, [0.0, 0.02, -0.7], 5, 0.3, 333, 201, 0.005),
    "box_r128": (lambda: sdf_box(128), 128, [-0.05, 0.0, -0.8], 7, 0.25, 320, 240, 0.004),
    "sphere_r2": (lambda: sdf_sphere(2, r=0.2), 2, [0.0, 0.0, -1.0], 9, 0.3, 64, 48, 0.01),
    "bowl_camera_inside_box": (lambda: sdf_bowl(48), 48, [0.0, 0.05, -0.1], 11, 0.5, 160, 120, 0.01),
    "bottle_partly_behind": (lambda: sdf_bottle(40), 40, [0.1, 0.0, -0.2], 13, 0.4, 160, 120, 0.005),
}


def scene(name):
    b, R, pos, qs, scale, W, H, thr = SCENES[name]
    sdf = np.ascontiguousarray(b(), dtype=np.float32)
    assert sdf.shape[0] == R
    return sdf, np.array(pos, np.float32), shoemake(qs), np.float32(1 / scale), W, H, thr


@pytest.mark.parametrize("name", list(SCENES))
def test_forward_matches_oracle(cuda_device, name):
    sdf, pos, q, inv_s, W, H, thr = scene(name)
    cam = default_camera(W, H)
    depth = render_depth_gpu(T(sdf, cuda_device), T(pos, cuda_device), T(q, cuda_device),
                             T([inv_s], cuda_device), threshold=thr, camera=cam_obj(W, H, cam))
    ref = oracle.render(sdf, pos, q, inv_s, W, H, threshold=thr, nthreads=8, **cam)
    info = depth_parity(depth.cpu().numpy(), ref, thr, rtol=DEPTH_RTOL)
    if name != "sphere_r2":
        assert info["n"] > 100, info


@pytest.mark.parametrize("name", ["mug_default_view", "torus_r32_odd_image", "box_r128",
                                  "bowl_camera_inside_box"])
@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_backward_matches_oracle(cuda_device, name, mode):
    sdf, pos, q, inv_s, W, H, thr = scene(name)
    cam = default_camera(W, H)
    tsdf, tp, tq = T(sdf, cuda_device, True), T(pos, cuda_device, True), T(q, cuda_device, True)
    ts = T([inv_s], cuda_device, True)
    depth = render_depth_gpu(tsdf, tp, tq, ts, threshold=thr, camera=cam_obj(W, H, cam),
                             sdf_grad_mode=mode)
    g = np.random.default_rng(1).standard_normal((H, W)).astype(np.float32)
    depth.backward(T(g, cuda_device))
    d_gpu = depth.detach().cpu().numpy()
    # feed the oracle the GPU's own depth so that a flipped pixel cannot leak into the gradient gate
    bw = oracle.render_backward(g, d_gpu, sdf, pos, q, inv_s, W, H, sdf_grad_mode=mode,
                                nthreads=8, **cam)
    grad_close(tp.grad.cpu().numpy(), bw["g_position"], GRAD_RTOL, "position")
    grad_close(tq.grad.cpu().numpy(), bw["g_orientation"], GRAD_RTOL, "orientation")
    grad_close(ts.grad.cpu().numpy(), [bw["g_inv_scale"]], GRAD_RTOL, "inv_scale")
    grad_close(tsdf.grad.cpu().numpy(), bw["g_sdf"], GRAD_RTOL, "sdf")


# ------------------------------------------------------------------------------------------
# (3) the reference's own CUDA extension on the same GPU
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("view", [0, 1, 2])
def test_against_reference_cuda_extension(cuda_device, ref_ext, view):
    sdf = mug_sdf()
    W, H = 640, 480
    cam = default_camera(W, H)
    pos = [[0.02, -0.01, -0.4], [0.1, 0.05, -0.9], [-0.05, 0.03, -0.25]][view]
    scale = [0.15, 0.055, 0.1][view]
    thr = [0.005, 0.003, 0.01][view]
    q = shoemake(20 + view)
    args = [T(sdf, cuda_device, True), T(pos, cuda_device, True), T(q, cuda_device, True),
            T([1 / scale], cuda_device, True)]
    depth = render_depth_gpu(*args, threshold=thr, camera=cam_obj(W, H, cam))
    with torch.no_grad():
        (ref_depth,) = ref_ext.forward(*[a.detach() for a in args], W, H, cam["cx"], cam["cy"],
                                       cam["fx"], cam["fy"], thr)
    info = depth_parity(depth.detach().cpu().numpy(), ref_depth.cpu().numpy(), thr,
                        rtol=DEPTH_RTOL)
    print("vs reference extension:", info)
    assert info["n"] > 1000

    g = torch.as_tensor(np.random.default_rng(view).standard_normal((H, W)).astype(np.float32),
                        device=cuda_device)
    depth.backward(g)
    # the reference backward on OUR depth image (identical inputs for the gradient comparison)
    r_sdf, r_p, r_q, r_s = ref_ext.backward(g, depth.detach(), *[a.detach() for a in args], W, H,
                                            cam["cx"], cam["cy"], cam["fx"], cam["fy"])
    grad_close(args[1].grad.cpu().numpy(), r_p.cpu().numpy(), GRAD_RTOL, "position vs ref ext")
    grad_close(args[2].grad.cpu().numpy(), r_q.cpu().numpy(), GRAD_RTOL, "orientation vs ref ext")
    grad_close(args[3].grad.cpu().numpy(), r_s.cpu().numpy(), GRAD_RTOL, "inv_scale vs ref ext")
    grad_close(args[0].grad.cpu().numpy(), r_sdf.cpu().numpy(), GRAD_RTOL, "sdf vs ref ext")


def test_oracle_f32_matches_reference_cuda_extension(cuda_device, ref_ext):
    """Pins the float32 oracle to the reference CUDA kernels themselves."""
    sdf = mug_sdf()
    W, H = 320, 240
    cam = default_camera(W, H)
    pos, q, inv_s, thr = np.array([0.02, -0.01, -0.4], np.float32), shoemake(1), 1 / 0.15, 0.005
    (ref_depth,) = ref_ext.forward(T(sdf, cuda_device), T(pos, cuda_device), T(q, cuda_device),
                                   T([inv_s], cuda_device), W, H, cam["cx"], cam["cy"],
                                   cam["fx"], cam["fy"], thr)
    mine = oracle.render(sdf, pos, q, inv_s, W, H, threshold=thr, nthreads=8, **cam)
    depth_parity(mine, ref_depth.cpu().numpy(), thr, rtol=DEPTH_RTOL)
    g = np.random.default_rng(3).standard_normal((H, W)).astype(np.float32)
    r_sdf, r_p, r_q, r_s = ref_ext.backward(T(g, cuda_device), ref_depth, T(sdf, cuda_device),
                                            T(pos, cuda_device), T(q, cuda_device),
                                            T([inv_s], cuda_device), W, H, cam["cx"], cam["cy"],
                                            cam["fx"], cam["fy"])
    bw = oracle.render_backward(g, ref_depth.cpu().numpy(), sdf, pos, q, inv_s, W, H,
                                sdf_grad_mode="reference", nthreads=8, **cam)
    grad_close(bw["g_position"], r_p.cpu().numpy(), GRAD_RTOL, "oracle position vs ref ext")
    grad_close(bw["g_orientation"], r_q.cpu().numpy(), GRAD_RTOL, "oracle orientation vs ref ext")
    grad_close([bw["g_inv_scale"]], r_s.cpu().numpy(), GRAD_RTOL, "oracle inv_scale vs ref ext")
    grad_close(bw["g_sdf"], r_sdf.cpu().numpy(), GRAD_RTOL, "oracle sdf vs ref ext")


# ------------------------------------------------------------------------------------------
# batched / fused / composite entry points
# ------------------------------------------------------------------------------------------
def hypotheses(B, seed=0):
    rng = np.random.default_rng(seed)
    pos = np.array([0.02, -0.01, -0.45], np.float32) + rng.normal(0, 0.03, (B, 3)).astype(np.float32)
    quat = np.stack([shoemake(seed * 1000 + i) for i in range(B)])
    inv_s = (1 / (0.15 * (1 + rng.uniform(-0.1, 0.1, B)))).astype(np.float32)
    return pos, quat, inv_s


def test_batched_equals_single_renders(cuda_device):
    B, W, H, thr = 5, 200, 150, 0.005
    cam = cam_obj(W, H, default_camera(W, H))
    grids = np.stack([mug_sdf() + np.float32(0.002 * i) for i in range(B)])
    pos, quat, inv_s = hypotheses(B)
    tg = T(grids, cuda_device, True)
    tp, tq, ts = T(pos, cuda_device, True), T(quat, cuda_device, True), T(inv_s, cuda_device, True)
    depth = render_depth_batched(tg, tp, tq, ts, thr, cam)
    g = torch.randn(B, H, W, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(0))
    depth.backward(g)
    for b in range(B):
        a = [T(grids[b], cuda_device, True), T(pos[b], cuda_device, True),
             T(quat[b], cuda_device, True), T(inv_s[b:b + 1], cuda_device, True)]
        d1 = render_depth_gpu(*a, threshold=thr, camera=cam)
        assert torch.equal(d1, depth[b].detach())
        d1.backward(g[b])
        grad_close(tp.grad[b].cpu().numpy(), a[1].grad.cpu().numpy(), 1e-4, "batched position")
        grad_close(tq.grad[b].cpu().numpy(), a[2].grad.cpu().numpy(), 1e-4, "batched orientation")
        grad_close(ts.grad[b:b + 1].cpu().numpy(), a[3].grad.cpu().numpy(), 1e-4, "batched scale")
        grad_close(tg.grad[b].cpu().numpy(), a[0].grad.cpu().numpy(), 1e-4, "batched sdf")


def test_shared_grid_accumulates_sdf_gradients(cuda_device):
    B, W, H, thr = 3, 160, 120, 0.005
    cam = cam_obj(W, H, default_camera(W, H))
    grid = mug_sdf()
    pos, quat, inv_s = hypotheses(B, seed=1)
    shared = T(grid, cuda_device, True)
    d = render_depth_batched(shared, T(pos, cuda_device), T(quat, cuda_device),
                             T(inv_s, cuda_device), thr, cam)
    g = torch.randn(B, H, W, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(1))
    d.backward(g)
    per = T(np.stack([grid] * B), cuda_device, True)
    d2 = render_depth_batched(per, T(pos, cuda_device), T(quat, cuda_device),
                              T(inv_s, cuda_device), thr, cam)
    assert torch.equal(d, d2)
    d2.backward(g)
    grad_close(shared.grad.cpu().numpy(), per.grad.sum(0).cpu().numpy(), 1e-4, "shared grid")


def test_fused_compare_equals_unfused_composition(cuda_device):
    """render_and_compare == render_depth_batched + the reference's masked L1
    (estimation/simple_setup.py:125-131) in torch, and both match the oracle's loss."""
    B, W, H, thr = 4, 320, 240, 0.005
    cam_d = default_camera(W, H)
    cam = cam_obj(W, H, cam_d)
    grids = np.stack([mug_sdf() + np.float32(0.001 * i) for i in range(B)])
    pos, quat, inv_s = hypotheses(B, seed=2)
    obs = oracle.render(mug_sdf(), [0.02, -0.01, -0.45], shoemake(2000), 1 / 0.15, W, H,
                        threshold=thr, nthreads=8, **cam_d)
    t_obs = T(obs, cuda_device)

    def leaves():
        return [T(grids, cuda_device, True), T(pos, cuda_device, True),
                T(quat, cuda_device, True), T(inv_s, cuda_device, True)]

    a = leaves()
    loss, depth, n = render_and_compare(*a, t_obs, thr, cam)
    w = torch.tensor([1.0, 0.5, 2.0, 1.5], device=cuda_device)
    (loss * w).sum().backward()

    b = leaves()
    d2 = render_depth_batched(*b, thr, cam)
    assert torch.equal(depth, d2.detach())
    mask = (t_obs > 0)[None] & (d2 > 0)
    err = torch.abs(d2 - t_obs[None])
    loss2 = torch.stack([err[i][mask[i]].mean() for i in range(B)])
    (loss2 * w).sum().backward()

    assert torch.equal(n, mask.sum((1, 2)).float())
    grad_close(loss.detach().cpu().numpy(), loss2.detach().cpu().numpy(), 1e-5, "fused loss")
    for i, nm in enumerate(["sdf", "position", "orientation", "inv_scale"]):
        grad_close(a[i].grad.cpu().numpy(), b[i].grad.cpu().numpy(), GRAD_RTOL, "fused " + nm)
    for i in range(B):
        lo, _, no = oracle.l1_depth_loss(depth[i].cpu().numpy(), obs)
        assert no == int(n[i].item())
        assert abs(lo - loss[i].item()) <= 1e-5 * lo


def test_compare_without_overlap_gives_nan_loss_and_zero_grads(cuda_device):
    W, H = 64, 48
    cam = cam_obj(W, H, default_camera(W, H))
    a = [T(sdf_sphere(16)[None], cuda_device, True), T([[0, 0, -1.0]], cuda_device, True),
         T([[0, 0, 0, 1.0]], cuda_device, True), T([2.5], cuda_device, True)]
    loss, depth, n = render_and_compare(*a, torch.zeros(H, W, device=cuda_device), 0.005, cam)
    assert n.item() == 0 and torch.isnan(loss).all() and (depth > 0).any()
    g = torch.autograd.grad(loss.sum(), a[1:], allow_unused=True)
    assert all(float(x.abs().max()) == 0 for x in g)


def test_composite_matches_oracle(cuda_device):
    K, R, W, H, thr = 6, 32, 320, 180, 0.005
    cam_d = dict(cx=W / 2, cy=H / 2, fx=W / 2, fy=W / 2)
    cam = cam_obj(W, H, cam_d)
    builders = [sdf_sphere, sdf_torus, sdf_box, sdf_bottle, sdf_bowl, sdf_sphere]
    grids = np.stack([b(R) for b in builders])
    pos = np.array([[-0.3, 0.1, -0.9], [0.0, 0.1, -0.8], [0.3, 0.1, -1.0],
                    [-0.15, -0.1, -0.6], [0.15, -0.1, -0.7], [0.05, 0.0, -1.2]], np.float32)
    quat = np.stack([shoemake(50 + k) for k in range(K)])
    inv_s = (1 / np.array([0.15, 0.2, 0.15, 0.2, 0.15, 0.4], np.float32)).astype(np.float32)
    a = [T(grids, cuda_device, True), T(pos, cuda_device, True), T(quat, cuda_device, True),
         T(inv_s, cuda_device, True)]
    depth, winner = render_depth_composite(*a, thr, cam)
    d_or, w_or = oracle.render_composite(grids, pos, quat, inv_s, W, H, threshold=thr,
                                         nthreads=8, **cam_d)
    depth_parity(depth.detach().cpu().numpy(), d_or, thr, rtol=DEPTH_RTOL)
    w_gpu = winner.cpu().numpy()
    assert (w_gpu == w_or).mean() > 0.999
    assert set(np.unique(w_gpu)) >= {-1, 0, 1, 2, 3, 4}
    g = np.random.default_rng(5).standard_normal((H, W)).astype(np.float32)
    depth.backward(T(g, cuda_device))
    bws = oracle.render_composite_backward(g, depth.detach().cpu().numpy(), w_gpu, grids, pos,
                                           quat, inv_s, W, H, nthreads=8, **cam_d)
    for k in range(K):
        if not (w_gpu == k).any():
            continue
        grad_close(a[1].grad[k].cpu().numpy(), bws[k]["g_position"], GRAD_RTOL, f"obj{k} pos")
        grad_close(a[2].grad[k].cpu().numpy(), bws[k]["g_orientation"], GRAD_RTOL, f"obj{k} quat")
        grad_close(a[3].grad[k:k + 1].cpu().numpy(), [bws[k]["g_inv_scale"]], GRAD_RTOL, f"obj{k} s")
        grad_close(a[0].grad[k].cpu().numpy(), bws[k]["g_sdf"], GRAD_RTOL, f"obj{k} sdf")


def test_composite_of_one_object_is_the_plain_render(cuda_device):
    W, H, thr = 160, 120, 0.005
    cam = cam_obj(W, H, default_camera(W, H))
    a = [T(mug_sdf()[None], cuda_device), T([[0.02, -0.01, -0.4]], cuda_device),
         T(shoemake(1)[None], cuda_device), T([1 / 0.15], cuda_device)]
    depth, winner = render_depth_composite(*a, thr, cam)
    plain = render_depth_batched(*a, thr, cam)[0]
    assert torch.equal(depth, plain)
    assert torch.equal(winner >= 0, plain > 0)


# ------------------------------------------------------------------------------------------
# reference calling conventions and edge cases
# ------------------------------------------------------------------------------------------
def test_reference_argument_shapes(cuda_device):
    """Callers pass (3,)/(4,)/0-dim (simple_setup.py:432-434) and (1,3)/(1,4)/(1,)
    (simple_setup.py:517-519, vae/scripts/train.py:255-265); grads come back in those shapes."""
    W, H, thr = 96, 72, 0.005
    cam = cam_obj(W, H, default_camera(W, H))
    sdf = T(mug_sdf(), cuda_device)
    base = None
    for pshape, qshape, sshape in [((3,), (4,), ()), ((1, 3), (1, 4), (1,)), ((3,), (1, 4), (1,))]:
        p = T(np.array([0.02, -0.01, -0.4]).reshape(pshape), cuda_device, True)
        q = T(shoemake(1).reshape(qshape), cuda_device, True)
        s = T(np.array(1 / 0.15).reshape(sshape), cuda_device, True)
        d = render_depth_gpu(sdf, p, q, s, None, None, None, thr, cam)
        d.sum().backward()
        assert p.grad.shape == pshape and q.grad.shape == qshape and s.grad.shape == sshape
        assert sdf.grad is None
        if base is None:
            base = d.detach()
        assert torch.equal(base, d.detach())


def test_only_requested_gradients_are_computed(cuda_device):
    W, H, thr = 96, 72, 0.005
    cam = cam_obj(W, H, default_camera(W, H))
    sdf = T(mug_sdf(), cuda_device, True)
    p = T([0.02, -0.01, -0.4], cuda_device)
    q = T(shoemake(1), cuda_device)
    s = T([1 / 0.15], cuda_device)
    render_depth_gpu(sdf, p, q, s, threshold=thr, camera=cam).sum().backward()
    assert sdf.grad is not None and float(sdf.grad.abs().sum()) > 0
    assert p.grad is None and q.grad is None and s.grad is None


def test_object_out_of_view_and_behind_camera(cuda_device):
    W, H = 64, 48
    cam = cam_obj(W, H, default_camera(W, H))
    sdf = T(sdf_sphere(16), cuda_device)
    for pos in ([5.0, 0, -1.0], [0, 0, 2.0], [0, -4.0, -0.5]):
        d = render_depth_gpu(sdf, T(pos, cuda_device), T([0, 0, 0, 1.0], cuda_device),
                             T([2.5], cuda_device), threshold=0.005, camera=cam)
        assert float(d.abs().max()) == 0.0


def test_depth_buffer_is_fully_written(cuda_device):
    """No reliance on pre-zeroed output (the reference needs torch::zeros, cu:484)."""
    lib = _lib.lib()
    W, H = 70, 50
    cam = default_camera(W, H)
    sdf, p = T(sdf_sphere(16), cuda_device), T([0.3, 0, -1.0], cuda_device)
    q, s = T([0, 0, 0, 1.0], cuda_device), T([4.0], cuda_device)
    out = torch.full((H, W), float("nan"), device=cuda_device)
    _lib.check(lib.sdfr_forward(sdf.data_ptr(), 16, 0, 0, p.data_ptr(), q.data_ptr(), s.data_ptr(), 1,
                                W, H, cam["cx"], cam["cy"], cam["fx"], cam["fy"], 0.005,
                                out.data_ptr(), None, torch.cuda.current_stream().cuda_stream), "fwd")
    assert not torch.isnan(out).any() and (out > 0).any() and (out == 0).any()


def test_zero_threshold_terminates(cuda_device):
    """threshold=0 is the Python default of the reference (sdf_renderer.py:277); its loop can
    spin forever on an exact-zero sample; ours is capped."""
    W, H = 64, 48
    cam = cam_obj(W, H, default_camera(W, H))
    sdf = T(np.zeros((8, 8, 8), np.float32), cuda_device)
    d = render_depth_gpu(sdf, T([0, 0, -1.0], cuda_device), T([0, 0, 0, 1.0], cuda_device),
                         T([2.0], cuda_device), threshold=0.0, camera=cam)
    torch.cuda.synchronize()
    assert float(d.abs().max()) == 0.0
    st = forward_stats(sdf[None], T([[0, 0, -1.0]], cuda_device), T([[0, 0, 0, 1.0]], cuda_device),
                       T([2.0], cuda_device), 0.0, cam)
    assert st["capped_rays"] == st["box_pixels"] > 0
    assert st["samples"] == st["capped_rays"] * _lib.lib().sdfr_max_steps()


def test_forward_stats_match_oracle_counts(cuda_device):
    sdf, pos, q, inv_s, W, H, thr = scene("mug_default_view")
    cam = default_camera(W, H)
    st = forward_stats(T(sdf[None], cuda_device), T(pos[None], cuda_device), T(q[None], cuda_device),
                       T([inv_s], cuda_device), thr, cam_obj(W, H, cam))
    d, steps, _ = oracle.render(sdf, pos, q, inv_s, W, H, threshold=thr, extras=True, nthreads=8,
                                **cam)
    assert abs(st["samples"] - int(steps.sum())) <= 2e-3 * steps.sum()
    assert abs(st["hit_pixels"] - int((d > 0).sum())) <= 5
    assert st["capped_rays"] == 0


def test_non_default_stream_and_graph_capture(cuda_device):
    """Launches go to the caller's stream and are CUDA-graph capturable (the reference uses the
    legacy default stream and allocates inside the call, cu:484-495)."""
    B, W, H, thr = 3, 160, 120, 0.005
    cam = cam_obj(W, H, default_camera(W, H))
    pos, quat, inv_s = hypotheses(B, seed=4)
    a = [T(mug_sdf(), cuda_device), T(pos, cuda_device), T(quat, cuda_device), T(inv_s, cuda_device)]
    want = render_depth_batched(*a, thr, cam)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        got = render_depth_batched(*a, thr, cam)
    side.synchronize()
    assert torch.equal(want, got)

    lib = _lib.lib()
    out = torch.zeros(B, H, W, device=cuda_device)
    cp = default_camera(W, H)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        _lib.check(lib.sdfr_forward(a[0].data_ptr(), 64, 0, 0, a[1].data_ptr(), a[2].data_ptr(),
                                    a[3].data_ptr(), B, W, H, cp["cx"], cp["cy"], cp["fx"],
                                    cp["fy"], thr, out.data_ptr(), None,
                                    torch.cuda.current_stream().cuda_stream), "captured forward")
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(want, out)


def test_backward_is_linear_in_grad_depth(cuda_device):
    """Size-independent property at the full reference size (640x480, 64^3)."""
    sdf, pos, q, inv_s, W, H, thr = scene("mug_default_view")
    cam = cam_obj(W, H, default_camera(W, H))
    gen = torch.Generator(cuda_device).manual_seed(7)
    g1 = torch.randn(H, W, device=cuda_device, generator=gen)
    g2 = torch.randn(H, W, device=cuda_device, generator=gen)

    def grads(g):
        a = [T(sdf, cuda_device, True), T(pos, cuda_device, True), T(q, cuda_device, True),
             T([inv_s], cuda_device, True)]
        render_depth_gpu(*a, threshold=thr, camera=cam).backward(g)
        return [x.grad for x in a]

    ga, gb, gc = grads(g1), grads(g2), grads(2.0 * g1 - 0.5 * g2)
    for x, y, zc in zip(ga, gb, gc):
        grad_close(zc.cpu().numpy(), (2.0 * x - 0.5 * y).cpu().numpy(), 1e-4, "linearity")


def test_translation_along_the_optical_axis_property(cuda_device):
    """A sphere rendered at two distances: centre-pixel depth differs by exactly the shift
    (up to the sphere-trace tolerance); checks the OpenGL sign conventions end to end."""
    W, H, thr = 65, 49, 0.001
    cam = cam_obj(W, H, dict(cx=32.5, cy=24.5, fx=60.0, fy=60.0))
    sdf = T(sdf_sphere(64, r=0.5), cuda_device)
    q, s = T([0, 0, 0, 1.0], cuda_device), T([1 / 0.3], cuda_device)
    d1 = render_depth_gpu(sdf, T([0, 0, -1.0], cuda_device), q, s, threshold=thr, camera=cam)
    d2 = render_depth_gpu(sdf, T([0, 0, -1.5], cuda_device), q, s, threshold=thr, camera=cam)
    c1, c2 = d1[24, 32].item(), d2[24, 32].item()
    assert abs(c1 - (1.0 - 0.15)) < 5e-3 and abs((c2 - c1) - 0.5) < 5e-3


def test_rejects_bad_inputs_on_gpu(cuda_device):
    cam = cam_obj(32, 24, default_camera(32, 24))
    good = [T(sdf_sphere(8), cuda_device), T([0, 0, -1.0], cuda_device),
            T([0, 0, 0, 1.0], cuda_device), T([2.0], cuda_device)]
    with pytest.raises(RuntimeError, match="float32"):
        render_depth_gpu(good[0].double(), *good[1:], threshold=0.01, camera=cam)
    with pytest.raises(RuntimeError, match="contiguous"):
        render_depth_gpu(good[0].transpose(0, 2), *good[1:], threshold=0.01, camera=cam)
    with pytest.raises(RuntimeError, match="at least 4"):
        render_depth_gpu(good[0], good[1], good[1], good[3], threshold=0.01, camera=cam)
    with pytest.raises(RuntimeError, match="shape"):
        render_depth_gpu(good[0][:4], *good[1:], threshold=0.01, camera=cam)
    assert render_depth_batched(good[0], torch.zeros(0, 3, device=cuda_device),
                                torch.zeros(0, 4, device=cuda_device),
                                torch.zeros(0, device=cuda_device), 0.01, cam).shape == (0, 24, 32)


# ------------------------------------------------------------------------------------------
# fused compare+backward traversal, deferred normalisation, and the batched loop
# ------------------------------------------------------------------------------------------
def _compare_setup(cuda_device, B=3, W=320, H=240, shared=False, seed=6):
    thr = 0.005
    cam_d = default_camera(W, H)
    cam = cam_obj(W, H, cam_d)
    grid = mug_sdf()
    grids = grid if shared else np.stack([grid + np.float32(0.001 * i) for i in range(B)])
    pos, quat, inv_s = hypotheses(B, seed=seed)
    obs = oracle.render(grid, [0.02, -0.01, -0.45], shoemake(seed * 1000), 1 / 0.15, W, H,
                        threshold=thr, nthreads=8, **cam_d)
    mk = lambda: [T(grids, cuda_device, True), T(pos, cuda_device, True),
                  T(quat, cuda_device, True), T(inv_s, cuda_device, True)]
    return mk, T(obs, cuda_device), thr, cam


def test_fused_traversal_equals_two_kernel_path(cuda_device):
    """sdfr_compare_fused + sdfr_scale_grads == sdfr_compare_forward + sdfr_compare_backward."""
    mk, obs, thr, cam = _compare_setup(cuda_device)
    w = torch.tensor([1.0, 0.3, 2.5], device=cuda_device)
    a = mk()
    loss, depth, n = render_and_compare(*a, obs, thr, cam)
    (loss * w).sum().backward(retain_graph=True)
    first = [x.grad.clone() for x in a]
    for x in a:
        x.grad = None
    # the second backward cannot reuse the (already scaled) fused buffers: it re-traverses with
    # the two-kernel path and must give the same gradients
    (loss * w).sum().backward()
    for nm, f, x in zip(["sdf", "position", "orientation", "inv_scale"], first, a):
        grad_close(x.grad.cpu().numpy(), f.cpu().numpy(), 2e-4, "fused vs two-kernel " + nm)
    assert all(float(f.abs().max()) > 0 for f in first)


def test_shared_grid_compare_falls_back_and_accumulates(cuda_device):
    mk, obs, thr, cam = _compare_setup(cuda_device, shared=True)
    a = mk()
    loss, depth, n = render_and_compare(*a, obs, thr, cam)
    loss.sum().backward()
    per = mk()
    per[0] = torch.stack([a[0].detach()] * 3).requires_grad_(True)
    loss2, _, _ = render_and_compare(*per, obs, thr, cam)
    loss2.sum().backward()
    # same pixels, but the per-warp partial sums reach the loss accumulator through float atomics
    # in launch-dependent order: equal to rounding, not bit for bit
    assert torch.allclose(loss, loss2, rtol=1e-5, atol=0)
    grad_close(a[0].grad.cpu().numpy(), per[0].grad.sum(0).cpu().numpy(), 2e-4, "shared compare sdf")
    grad_close(a[1].grad.cpu().numpy(), per[1].grad.cpu().numpy(), 2e-4, "shared compare pos")


def test_compare_only_pose_or_only_sdf_gradients(cuda_device):
    mk, obs, thr, cam = _compare_setup(cuda_device)
    full = mk()
    render_and_compare(*full, obs, thr, cam)[0].sum().backward()
    only_sdf = mk()
    for x in only_sdf[1:]:
        x.requires_grad_(False)
    render_and_compare(*only_sdf, obs, thr, cam)[0].sum().backward()
    grad_close(only_sdf[0].grad.cpu().numpy(), full[0].grad.cpu().numpy(), 1e-4, "sdf only")
    only_pose = mk()
    only_pose[0].requires_grad_(False)
    render_and_compare(*only_pose, obs, thr, cam)[0].sum().backward()
    for i in (1, 2, 3):
        grad_close(only_pose[i].grad.cpu().numpy(), full[i].grad.cpu().numpy(), 1e-4, "pose only")
    with torch.no_grad():
        l, d, n = render_and_compare(*mk(), obs, thr, cam)
    assert torch.isfinite(l).all() and (n > 0).all()


def test_single_frame_small_batch_paths(cuda_device):
    """B=1 uses one CTA per box tile with tile-local ray tables; must equal the oracle too."""
    for W, H in ((640, 480), (100, 37)):
        cam_d = default_camera(W, H)
        sdf, pos, q, inv_s = mug_sdf(), np.float32([0.02, -0.01, -0.4]), shoemake(1), np.float32(1 / 0.15)
        d = render_depth_gpu(T(sdf, cuda_device), T(pos, cuda_device), T(q, cuda_device),
                             T([inv_s], cuda_device), threshold=0.005, camera=cam_obj(W, H, cam_d))
        ref = oracle.render(sdf, pos, q, inv_s, W, H, threshold=0.005, nthreads=8, **cam_d)
        depth_parity(d.cpu().numpy(), ref, 0.005, rtol=DEPTH_RTOL)


def test_large_batch_few_ctas_per_hypothesis(cuda_device):
    """B large enough that every CTA walks many tiles (G small): same images as one by one."""
    B, W, H, thr = 300, 96, 64, 0.005
    cam = cam_obj(W, H, default_camera(W, H))
    pos, quat, inv_s = hypotheses(B, seed=9)
    a = [T(mug_sdf(), cuda_device), T(pos, cuda_device), T(quat, cuda_device), T(inv_s, cuda_device)]
    d = render_depth_batched(*a, thr, cam)
    for b in (0, 17, 299):
        one = render_depth_gpu(a[0], a[1][b], a[2][b], a[3][b:b + 1], threshold=thr, camera=cam)
        assert torch.equal(one, d[b])


def test_hypothesis_optimizer_recovers_pose(cuda_device):
    """The batched render-and-compare loop (simple_setup.py:400-470 for B hypotheses): losses
    fall and the best hypothesis approaches the pose that generated the observation."""
    from sdfest_b200.estimation import HypothesisOptimizer
    from sdfest_b200 import synthetic as syn

    W, H, thr, B = 320, 240, 0.005, 8
    cam = cam_obj(W, H, default_camera(W, H))
    grid = syn.sdf_mug(64, cuda_device)
    hyp = syn.make_hypotheses(B, seed=3, device=cuda_device, pos_sigma=0.01, rot_deg=5.0,
                              scale_rel=0.05)
    true_p, true_q = hyp["position"][0:1].clone(), hyp["orientation"][0:1].clone()
    true_s = 1.0 / hyp["inv_scale"][0:1]
    obs = render_depth_batched(grid, true_p, true_q, 1.0 / true_s, thr, cam)[0].contiguous()
    opt = HypothesisOptimizer(cam, thr, obs, hyp["position"][1:], hyp["orientation"][1:],
                              1.0 / hyp["inv_scale"][1:], sdf=grid[None], max_points=2000)
    first = opt.step().clone()
    last = opt.run(60)
    assert torch.isfinite(last).all()
    assert (last < first).float().mean() > 0.8
    best = int(torch.argmin(last))
    err0 = torch.linalg.norm(hyp["position"][1:][best] - true_p[0]).item()
    err1 = torch.linalg.norm(opt.position[best].detach() - true_p[0]).item()
    assert err1 < max(0.5 * err0, 2e-3), (err0, err1)


def test_skewed_layout_gives_identical_results(cuda_device):
    """The pitched (bank-conflict-free) copy holds the same fp32 values: depth, loss and all four
    gradients must be bit-identical between the two layouts (only float-atomic order differs for
    the sums, hence allclose there), for compile-time (64, 32) and run-time (24) resolutions."""
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.differentiable_renderer import get_sdf_layout_policy, set_sdf_layout_policy

    old = get_sdf_layout_policy()
    try:
        for R in (64, 32, 24):
            B, W, H, thr = 3, 160, 120, 0.005
            cam = Camera(W, H, W / 2, W / 2, W / 2, H / 2, pixel_center=0.5)
            hyp = syn.make_hypotheses(B, seed=3, device=cuda_device)
            grids = syn.hypothesis_grids(hyp["shape_param"], R, cuda_device)
            res = {}
            for policy in ("dense", "skewed"):
                set_sdf_layout_policy(policy)
                a = [grids.clone().requires_grad_(True), hyp["position"].clone().requires_grad_(True),
                     hyp["orientation"].clone().requires_grad_(True),
                     hyp["inv_scale"].clone().requires_grad_(True)]
                d = render_depth_batched(*a, thr, cam)
                obs = d[0].detach().clone()
                loss, depth, n = render_and_compare(*a, obs, thr, cam)
                loss[1:].sum().backward()
                res[policy] = (d.detach(), depth, n, loss.detach(), [x.grad for x in a])
                assert (d > 0).sum() > 500
            dd, ds = res["dense"], res["skewed"]
            assert torch.equal(dd[0], ds[0]) and torch.equal(dd[1], ds[1]) and torch.equal(dd[2], ds[2])
            assert torch.allclose(dd[3][1:], ds[3][1:], rtol=1e-5, atol=0)
            for gd, gs in zip(dd[4], ds[4]):
                scale = float(gd.abs().max())
                assert scale > 0 and float((gd - gs).abs().max()) <= 2e-4 * scale
        # single-frame reference API through the skewed copy
        set_sdf_layout_policy("skewed")
        z = load_golden("mug_z0_r64")
        args = [torch.as_tensor(np.asarray(z[k], np.float32), device=cuda_device)
                for k in ("sdf", "position", "orientation")]
        inv = torch.tensor([float(z["inv_scale"])], device=cuda_device)
        cam = Camera(z["W"], z["H"], z["cam"]["fx"], z["cam"]["fy"], z["cam"]["cx"], z["cam"]["cy"],
                     pixel_center=0.5)
        d_sk = render_depth_gpu(*args, inv, threshold=float(z["threshold"]), camera=cam)
        set_sdf_layout_policy("dense")
        d_de = render_depth_gpu(*args, inv, threshold=float(z["threshold"]), camera=cam)
        assert torch.equal(d_sk, d_de)
    finally:
        set_sdf_layout_policy(old)


@pytest.mark.parametrize("graph", [False, True])
def test_streamed_host_buffers_match_device_path(cuda_device, graph):
    """estimation.StreamedRenderCompare (pinned host inputs, chunked 3-stream pipeline, optional
    CUDA-graph replay) returns the losses and gradients of render_and_compare + autograd."""
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.estimation import StreamedRenderCompare

    B, R, W, H, thr = 5, 32, 160, 120, 0.005
    cam = Camera(W, H, W / 2, W / 2, W / 2, H / 2, pixel_center=0.5)
    hyp = syn.make_hypotheses(B, seed=5, device=cuda_device)
    grids = syn.hypothesis_grids(hyp["shape_param"], R, cuda_device)
    a = [grids.clone().requires_grad_(True), hyp["position"].clone().requires_grad_(True),
         hyp["orientation"].clone().requires_grad_(True), hyp["inv_scale"].clone().requires_grad_(True)]
    obs = render_depth_batched(grids[:1], hyp["position"][:1] + 0.01, hyp["orientation"][:1],
                               hyp["inv_scale"][:1], thr, cam)[0].contiguous()
    loss, depth, n = render_and_compare(*a, obs, thr, cam)
    loss.sum().backward()
    host = [t.detach().cpu().pin_memory() for t in (grids, hyp["position"], hyp["orientation"],
                                                    hyp["inv_scale"], obs)]
    for sdf_to_host in (False, True):
        ev = StreamedRenderCompare(cam, thr, B, R, cuda_device, chunk=2, sdf_grads_to_host=sdf_to_host)
        for _ in range(2):  # second call replays the captured graph
            out = ev(*host, graph=graph)
        assert torch.allclose(out["loss"], loss.detach().cpu(), rtol=1e-5, atol=0)
        assert torch.equal(out["n_overlap"], n.cpu())
        assert torch.equal(out["depth_device"], depth)
        for got, want in ((out["g_position"], a[1].grad), (out["g_orientation"], a[2].grad),
                          (out["g_inv_scale"], a[3].grad), (out["g_sdf_device"], a[0].grad)):
            want = want.cpu()
            assert float((got.cpu() - want).abs().max()) <= 2e-4 * float(want.abs().max())
        if sdf_to_host:
            assert torch.equal(out["g_sdf_host"].view(B, R, R, R), out["g_sdf_device"].cpu())


def _pcloss_reference(points, pos, quat, scale, sdf):
    """float64 torch restatement of the reference's pc_loss (pinned to its golden vector by
    tests/test_estimation_host.py) -> per-hypothesis mean |.| and autograd gradients."""
    from oracle.pc_loss import pc_loss

    a = [t.detach().double().cpu().requires_grad_(True) for t in (pos, quat, scale, sdf)]
    pts = points.detach().double().cpu()
    if pts.dim() == 2:
        val = pc_loss(pts, *a).abs().mean(dim=1)
    else:
        val = torch.stack([pc_loss(pts[b], a[0][b:b + 1], a[1][b:b + 1], a[2][b:b + 1],
                                   a[3][b:b + 1] if a[3].shape[0] > 1 else a[3])[0].abs().mean()
                           for b in range(pts.shape[0])])
    return val, a


@pytest.mark.parametrize("shared_cloud", [True, False])
@pytest.mark.parametrize("shared_grid", [True, False])
def test_point_loss_kernel_matches_reference_formula(cuda_device, shared_cloud, shared_grid):
    """sdfr_point_loss_forward/backward against estimation/losses.py:32-135 (values 1e-5,
    gradients 1e-3 relative), un-normalised quaternions, points outside the grid included."""
    from sdfest_b200.estimation import point_loss

    g = torch.Generator().manual_seed(7)
    B, M, R = 4, 3000, 24
    sdf = torch.as_tensor(np.stack([sdf_torus(R), sdf_sphere(R), sdf_box(R), sdf_torus(R) * 0.7]),
                          dtype=torch.float32)
    if shared_grid:
        sdf = sdf[:1]
    pos = torch.tensor([0.05, -0.03, -0.8]) + 0.02 * torch.randn(B, 3, generator=g)
    quat = torch.as_tensor(np.stack([shoemake(20 + i) for i in range(B)]), dtype=torch.float32)
    quat = quat * (0.5 + torch.rand(B, 1, generator=g))  # un-normalised on purpose
    scale = 0.3 * (1 + 0.2 * (torch.rand(B, generator=g) - 0.5))
    pts = torch.tensor([0.05, -0.03, -0.8]) + (torch.rand(M if shared_cloud else B * M, 3, generator=g) - 0.5) * 0.8
    if not shared_cloud:
        pts = pts.view(B, M, 3)
    w = torch.tensor([1.0, 0.4, 2.0, 1.3])
    ref, a = _pcloss_reference(pts, pos, quat, scale, sdf)
    (ref * w.double()).sum().backward()
    d = [t.to(cuda_device).requires_grad_(True) for t in (pos, quat, scale, sdf)]
    got = point_loss(pts.to(cuda_device), *d)
    (got * w.to(cuda_device)).sum().backward()
    assert np.abs(got.detach().cpu().numpy() - ref.detach().numpy()).max() <= 1e-5 * ref.max().item()
    assert 0.05 < float((ref > 0).float().mean()) <= 1.0
    for nm, x, y in zip(("position", "orientation", "scale", "sdf"), d, a):
        grad_close(x.grad.cpu().numpy(), y.grad.numpy(), GRAD_RTOL, "point loss " + nm)
        assert float(y.grad.abs().max()) > 0


def test_point_loss_kernel_matches_reference_golden(cuda_device):
    """The reference's own pc_loss outputs and autograd gradients (tests/golden/pcloss_torus16)."""
    import os

    from sdfest_b200.estimation import point_loss
    from util import GOLDEN_DIR

    z = np.load(os.path.join(GOLDEN_DIR, "pcloss_torus16.npz"))
    t = lambda k: torch.tensor(np.asarray(z[k], np.float32), device=cuda_device)  # noqa: E731
    pos, quat = t("position")[None].requires_grad_(True), t("orientation")[None].requires_grad_(True)
    scale, sdf = t("scale").reshape(1).requires_grad_(True), t("sdf")[None].requires_grad_(True)
    got = point_loss(t("points"), pos, quat, scale, sdf)
    got.sum().backward()
    want = np.abs(z["value"]).mean()
    assert abs(got.item() - want) <= 1e-5 * want
    for x, key in ((pos, "g_position"), (quat, "g_orientation"), (scale, "g_scale"), (sdf, "g_sdf")):
        grad_close(x.grad.cpu().numpy().reshape(z[key].shape), z[key], GRAD_RTOL, key)
