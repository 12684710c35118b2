"""Decoder tail: final trilinear interpolation + 1x1x1 convolution of the
reference decoder (sdfest/vae/sdf_vae.py:235-247).

CPU part: the numpy oracle (oracle/decoder_tail.py) is pinned against vectors recorded from the
reference's own SDFDecoder (tests/golden/decoder_tail_*.npz) and against torch's CPU operators; the
host-side wrapper splits a decoder into trunk + tail correctly.  GPU part: the sm_100a kernels
(through the C ABI) against the oracle, the golden vectors and torch's CUDA operators.
Floating point, reassociated sums: tolerance 1e-5 of the largest magnitude (forward) and 1e-4
(backward, ~125-term sums)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import decoder_tail as odt
from sdfest_b200 import _lib
from sdfest_b200.estimation import FusedTailDecoder, SDFDecoder, decoder_tail

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD_TOL, BWD_TOL = 1e-5, 1e-4


def rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def torch_tail(x, w, b, R):
    y = F.interpolate(x, size=(R, R, R), mode="trilinear", align_corners=False)
    return F.conv3d(y, w.view(1, -1, 1, 1, 1), b.view(1) if b is not None else None)[:, 0]


# ---------------------------------------------------------------------------------------------
# CPU: oracle pinned to the reference decoder and to torch
# ---------------------------------------------------------------------------------------------
def test_oracle_matches_reference_decoder_golden_small():
    z = np.load(os.path.join(GOLDEN, "decoder_tail_small.npz"))
    assert rel(odt.tail_forward(z["x"], z["weight"], z["bias"], 12), z["out"]) < 1e-6
    assert rel(odt.tail_backward(z["g"], z["weight"], 5), z["g_x"]) < 1e-6


def test_oracle_matches_reference_mug_decoder_golden():
    m = np.load(os.path.join(GOLDEN, "decoder_tail_mug_z0.npz"))
    ref = np.load(os.path.join(GOLDEN, "mug_z0_sdf.npz"))["sdf"]  # the reference's decode(0)
    assert rel(odt.tail_forward(m["x"][None], m["weight"], m["bias"], 64)[0], ref) < 1e-6
    g = np.random.default_rng(int(m["g_seed"])).standard_normal((1, 1, 64, 64, 64)).astype(np.float32)[:, 0]
    gx = odt.tail_backward(g, m["weight"], 30)[0]
    assert rel(gx[:, m["g_x_planes"]], m["g_x"]) < 1e-6


@pytest.mark.parametrize("C,S,R", [(4, 30, 64), (3, 5, 12), (2, 8, 8), (1, 16, 6), (5, 1, 4)])
def test_oracle_matches_torch_cpu_operators(C, S, R):
    rng = np.random.default_rng(S * 100 + R)
    x = torch.tensor(rng.standard_normal((2, C, S, S, S)), dtype=torch.float32, requires_grad=True)
    w = torch.tensor(rng.standard_normal(C), dtype=torch.float32)
    b = torch.tensor(rng.standard_normal(()), dtype=torch.float32)
    out = torch_tail(x, w, b, R)
    g = torch.tensor(rng.standard_normal((2, R, R, R)), dtype=torch.float32)
    out.backward(g)
    assert rel(odt.tail_forward(x.detach().numpy(), w.numpy(), b.numpy(), R), out.detach().numpy()) < 1e-6
    assert rel(odt.tail_backward(g.numpy(), w.numpy(), S), x.grad.numpy()) < 1e-6
    coef = np.array([0.5, -2.0])
    extra = rng.standard_normal((2, R, R, R))
    want = odt.tail_backward(g.numpy() * coef[:, None, None, None] + extra, w.numpy(), S)
    assert rel(odt.tail_backward(g.numpy(), w.numpy(), S, coef=coef, g_extra=extra), want) < 1e-6


def test_axis_weights_are_a_partition_of_unity():
    for S, R in ((30, 64), (5, 12), (16, 6), (1, 4), (64, 64)):
        W = odt.axis_weights(S, R)
        assert W.shape == (R, S) and np.allclose(W.sum(axis=1), 1.0, atol=1e-6) and (W >= 0).all()
    assert np.array_equal(odt.axis_weights(7, 7), np.eye(7))


def test_wrapper_splits_trunk_and_tail_like_the_plain_decoder():
    torch.manual_seed(0)
    dec = SDFDecoder(64).eval()
    fused = FusedTailDecoder(dec)
    z = torch.randn(2, 8)
    x = fused.trunk(z)
    assert tuple(x.shape) == (2, 4, 30, 30, 30)
    w, b = fused.tail_parameters()
    with torch.no_grad():
        assert torch.allclose(torch_tail(x, w, b, 64), dec(z)[:, 0], atol=1e-6)
    assert all(not p.requires_grad for p in dec.parameters())


def test_wrapper_rejects_decoders_without_a_fusable_tail():
    conv = ((8, 16, 16, 3, True), (16, 16, 1, 3, False))  # last stage is a 3x3x3 convolution
    with pytest.raises(ValueError, match="last stage"):
        FusedTailDecoder(SDFDecoder(14, conv=conv))
    conv = ((8, 16, 4, 3, True), (16, 4, 1, 1, False))  # 16 != volume size 32: extra interpolation
    with pytest.raises(ValueError, match="last stage"):
        FusedTailDecoder(SDFDecoder(32, conv=conv))


def test_tail_argument_errors_need_no_gpu():
    lib = _lib.lib()
    assert lib.sdfr_decoder_tail_forward(None, 4, 30, None, None, None, 0, 64, None, 0, 0, None) == 0
    assert lib.sdfr_decoder_tail_forward(None, 4, 30, None, None, None, 1, 64, None, 64 ** 3, 0, None) == -1
    assert lib.sdfr_decoder_tail_forward(None, 17, 30, None, None, None, 1, 64, None, 64 ** 3, 0, None) == -2
    assert lib.sdfr_decoder_tail_forward(None, 4, 30, None, None, None, 1, 256, None, 256 ** 3, 0, None) == -2
    assert lib.sdfr_decoder_tail_forward(None, 4, 30, None, None, None, 1, 64, None, 64 ** 3, 5, None) == -3
    assert lib.sdfr_decoder_tail_backward(None, 0, None, None, None, 0, None, 4, 30, 0, 64, None, None) == 0
    assert lib.sdfr_decoder_tail_backward(None, 0, None, None, None, 0, None, 4, 30, 1, 64, None, None) == -1
    assert lib.sdfr_decoder_tail_backward(None, -1, None, None, None, 0, None, 4, 30, 1, 64, None, None) == -2
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        decoder_tail(torch.zeros(1, 4, 3, 3, 3), torch.zeros(4), None, 8)


# ---------------------------------------------------------------------------------------------
# GPU: kernels against oracle, golden vectors and torch's CUDA operators
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_tail_kernels_match_reference_golden(cuda_device):
    z = np.load(os.path.join(GOLDEN, "decoder_tail_small.npz"))
    x = torch.tensor(z["x"], device=cuda_device, requires_grad=True)
    w, b = torch.tensor(z["weight"], device=cuda_device), torch.tensor(z["bias"], device=cuda_device).view(1)
    out = decoder_tail(x, w, b, 12)
    assert rel(out.detach().cpu().numpy(), z["out"]) < FWD_TOL
    out.backward(torch.tensor(z["g"], device=cuda_device))
    assert rel(x.grad.cpu().numpy(), z["g_x"]) < BWD_TOL

    m = np.load(os.path.join(GOLDEN, "decoder_tail_mug_z0.npz"))
    ref = np.load(os.path.join(GOLDEN, "mug_z0_sdf.npz"))["sdf"]
    x = torch.tensor(m["x"][None], device=cuda_device, requires_grad=True)
    w, b = torch.tensor(m["weight"], device=cuda_device), torch.tensor(m["bias"], device=cuda_device).view(1)
    out = decoder_tail(x, w, b, 64)
    assert rel(out.detach().cpu().numpy()[0], ref) < FWD_TOL
    g = np.random.default_rng(int(m["g_seed"])).standard_normal((1, 1, 64, 64, 64)).astype(np.float32)[:, 0]
    out.backward(torch.tensor(g, device=cuda_device))
    assert rel(x.grad.cpu().numpy()[0][:, m["g_x_planes"]], m["g_x"]) < BWD_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("C,S,R,B", [(4, 30, 64, 3), (3, 5, 12, 2), (16, 8, 8, 1), (1, 16, 6, 2),
                                     (5, 1, 4, 1), (2, 60, 128, 1), (4, 128, 32, 1)])
def test_tail_kernels_match_oracle_and_torch(cuda_device, C, S, R, B):
    rng = np.random.default_rng(C + S + R)
    x = torch.tensor(rng.standard_normal((B, C, S, S, S)), dtype=torch.float32, device=cuda_device,
                     requires_grad=True)
    w = torch.tensor(rng.standard_normal(C), dtype=torch.float32, device=cuda_device)
    b = torch.tensor(rng.standard_normal(1), dtype=torch.float32, device=cuda_device)
    g = torch.tensor(rng.standard_normal((B, R, R, R)), dtype=torch.float32, device=cuda_device)
    out = decoder_tail(x, w, b, R)
    out.backward(g)
    got_out, got_gx = out.detach().cpu().numpy(), x.grad.cpu().numpy()
    xn, wn = x.detach().cpu().numpy(), w.cpu().numpy()
    assert rel(got_out, odt.tail_forward(xn, wn, float(b.item()), R)) < FWD_TOL
    assert rel(got_gx, odt.tail_backward(g.cpu().numpy(), wn, S)) < BWD_TOL
    # torch's own CUDA operators on the same inputs (cuDNN's default TF32 convolutions are ~1e-3)
    x2 = x.detach().clone().requires_grad_(True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch_tail(x2, w, b, R)
        ref.backward(g)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert rel(got_out, ref.detach().cpu().numpy().astype(np.float64)) < FWD_TOL
    assert rel(got_gx, x2.grad.cpu().numpy().astype(np.float64)) < BWD_TOL
    # no bias
    assert rel(decoder_tail(x.detach(), w, None, R).cpu().numpy(),
               odt.tail_forward(xn, wn, None, R)) < FWD_TOL


@pytest.mark.gpu
def test_tail_skewed_output_and_deferred_scaling(cuda_device):
    """Skewed output layout holds the same values; the backward folds coef = upstream/n_overlap
    and a second gradient grid into its load."""
    import ctypes

    lib = _lib.lib()
    C, S, R, B = 4, 30, 64, 3
    rng = np.random.default_rng(3)
    x = torch.tensor(rng.standard_normal((B, C, S, S, S)), dtype=torch.float32, device=cuda_device)
    w = torch.tensor(rng.standard_normal(C), dtype=torch.float32, device=cuda_device)
    b = torch.tensor([0.25], device=cuda_device)
    py, px, n = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_longlong(0)
    _lib.check(lib.sdfr_skewed_pitches(R, ctypes.byref(py), ctypes.byref(px), ctypes.byref(n)), "pitches")
    sk = torch.full((B, n.value), float("nan"), device=cuda_device)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.sdfr_decoder_tail_forward(x.data_ptr(), C, S, w.data_ptr(), b.data_ptr(), None, B, R,
                                             sk.data_ptr(), n.value, _lib.LAYOUT_SKEWED, st), "tail fwd")
    dense = decoder_tail(x, w, b, R)
    ix, iy, iz = torch.meshgrid(*(torch.arange(R, device=cuda_device),) * 3, indexing="ij")
    assert torch.equal(sk[:, (ix * px.value + iy * py.value + iz).reshape(-1)].view(B, R, R, R), dense)
    # a too-small stride for the layout is an argument error
    assert lib.sdfr_decoder_tail_forward(x.data_ptr(), C, S, w.data_ptr(), b.data_ptr(), None, B, R,
                                         sk.data_ptr(), R ** 3, _lib.LAYOUT_SKEWED, st) == -2

    g = torch.tensor(rng.standard_normal((B, R, R, R)), dtype=torch.float32, device=cuda_device)
    extra = torch.tensor(rng.standard_normal((B, R, R, R)), dtype=torch.float32, device=cuda_device)
    n_ov = torch.tensor([4.0, 0.0, 10.0], device=cuda_device)
    up = torch.tensor([2.0, 3.0, -1.0], device=cuda_device)
    gx = torch.empty(B, C, S, S, S, device=cuda_device)
    _lib.check(lib.sdfr_decoder_tail_backward(g.data_ptr(), R ** 3, n_ov.data_ptr(), up.data_ptr(),
                                              extra.data_ptr(), R ** 3, w.data_ptr(), C, S, B, R,
                                              gx.data_ptr(), st), "tail bwd")
    coef = np.array([0.5, 0.0, -0.1])
    want = odt.tail_backward(g.cpu().numpy(), w.cpu().numpy(), S, coef=coef, g_extra=extra.cpu().numpy())
    assert rel(gx.cpu().numpy(), want) < BWD_TOL
    # n_overlap == 0 with a NaN/inf-free result even if the raw gradient holds garbage there
    g[1] = float("inf")
    _lib.check(lib.sdfr_decoder_tail_backward(g.data_ptr(), R ** 3, n_ov.data_ptr(), up.data_ptr(),
                                              None, 0, w.data_ptr(), C, S, B, R, gx.data_ptr(), st), "tail bwd")
    assert torch.isfinite(gx).all() and float(gx[1].abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("C,S,R,layout", [(4, 30, 64, "skewed"), (4, 30, 64, "dense"), (3, 7, 18, "dense")])
def test_tail_forward_with_bounds_equals_tail_then_scan(cuda_device, C, S, R, layout):
    """sdfr_decoder_tail_forward_bounds writes the same grids as sdfr_decoder_tail_forward and the bounds
    sdfr_grid_bounds finds on them (both kernel paths: 4-row groups and the generic one)."""
    import ctypes

    lib, dev = _lib.lib(), cuda_device
    B = 5
    rng = np.random.default_rng(11)
    x = torch.tensor(rng.standard_normal((B, C, S, S, S)), dtype=torch.float32, device=dev)
    w = torch.tensor(rng.standard_normal(C) * 0.2, dtype=torch.float32, device=dev)
    b = torch.tensor([0.3], device=dev)
    # a base with a surface: a ball, so that the bounds are a proper sub-box for some hypotheses
    ax = torch.linspace(-1, 1, R, device=dev)
    base = ((ax[:, None, None] ** 2 + ax[None, :, None] ** 2 * 2 + ax[None, None, :] ** 2 * 4).sqrt() - 0.8).contiguous()
    pos = torch.tensor(rng.standard_normal((B, 3)) * 0.1 + [0, 0, -0.8], dtype=torch.float32, device=dev)
    inv_s = torch.tensor(1.0 / (0.1 + 0.2 * rng.random(B)), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    if layout == "skewed":
        n = ctypes.c_longlong(0)
        _lib.check(lib.sdfr_skewed_pitches(R, None, None, ctypes.byref(n)), "pitches")
        stride, lay = int(n.value), _lib.LAYOUT_SKEWED
    else:
        stride, lay = R ** 3, _lib.LAYOUT_DENSE
    for thr in (0.005, 0.2):
        g0 = torch.full((B, stride), -5.0, device=dev)
        g1 = torch.full((B, stride), -5.0, device=dev)
        b0 = torch.full((B, 8), -7, dtype=torch.int32, device=dev)
        b1 = torch.full((B, 8), -9, dtype=torch.int32, device=dev)
        _lib.check(lib.sdfr_decoder_tail_forward(x.data_ptr(), C, S, w.data_ptr(), b.data_ptr(), base.data_ptr(), B, R,
                                                 g0.data_ptr(), stride, lay, st), "tail")
        _lib.check(lib.sdfr_grid_bounds(g0.data_ptr(), R, stride, lay, pos.data_ptr(), inv_s.data_ptr(), B, thr,
                                        b0.data_ptr(), st), "scan")
        _lib.check(lib.sdfr_decoder_tail_forward_bounds(
            x.data_ptr(), C, S, w.data_ptr(), b.data_ptr(), base.data_ptr(), B, R, g1.data_ptr(), stride, lay,
            pos.data_ptr(), inv_s.data_ptr(), thr, b1.data_ptr(), st), "tail + bounds")
        torch.cuda.synchronize()
        assert torch.equal(g0, g1)
        assert torch.equal(b0[:, :7], b1[:, :7]), (b0, b1)
        assert bool((b0[:, 3] >= b0[:, 0]).any())  # not all empty
    assert lib.sdfr_decoder_tail_forward_bounds(x.data_ptr(), C, S, w.data_ptr(), None, None, B, R, g1.data_ptr(),
                                                stride, lay, None, inv_s.data_ptr(), 0.005, b1.data_ptr(), st) == -1


@pytest.mark.gpu
def test_fused_tail_decoder_matches_plain_decoder(cuda_device):
    torch.manual_seed(1)
    dec = SDFDecoder(64).to(cuda_device).eval()
    fused = FusedTailDecoder(dec)
    z1 = torch.randn(3, 8, device=cuda_device, requires_grad=True)
    z2 = z1.detach().clone().requires_grad_(True)
    g = torch.randn(3, 1, 64, 64, 64, device=cuda_device)
    torch.backends.cudnn.allow_tf32 = False  # the plain decoder's convolutions, for the comparison
    a, b = fused(z1), dec(z2)
    assert tuple(a.shape) == (3, 1, 64, 64, 64)
    assert rel(a.detach().cpu().numpy(), b.detach().cpu().numpy().astype(np.float64)) < FWD_TOL
    a.backward(g)
    b.backward(g)
    assert rel(z1.grad.cpu().numpy(), z2.grad.cpu().numpy().astype(np.float64)) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("with_points", [True, False])
def test_decode_render_compare_equals_the_composition_of_its_parts(cuda_device, with_points):
    """estimation.decode_render_compare (tail -> skewed grids -> fused compare -> point loss, one
    chain of C-ABI kernels) against the same loss assembled from the separate autograd operators."""
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.differentiable_renderer import Camera, render_and_compare, render_depth_batched
    from sdfest_b200.estimation import decode_render_compare, losses

    dev = cuda_device
    torch.manual_seed(2)
    B, R, W, H, thr = 3, 64, 160, 120, 0.005
    cam = Camera(W, H, 80.0, 80.0, 80.0, 60.0, pixel_center=0.5)
    hyp = syn.make_hypotheses(B, seed=3, device=dev)
    base = syn.sdf_mug(R, dev)
    x0 = 0.05 * torch.randn(B, 4, 30, 30, 30, device=dev)
    w = torch.randn(4, device=dev)
    bias = torch.tensor([0.01], device=dev)
    obs = render_depth_batched(base, hyp["position"][:1], hyp["orientation"][:1], hyp["inv_scale"][:1],
                               thr, cam)[0].contiguous()
    pts = losses.depth_to_pointcloud(obs, cam).contiguous() if with_points else None

    def leaves():
        return (x0.clone().requires_grad_(True), hyp["position"].clone().requires_grad_(True),
                hyp["orientation"].clone().requires_grad_(True),
                (1.0 / hyp["inv_scale"]).clone().requires_grad_(True))

    x, p, q, s = leaves()
    loss, depth, n, loss_d = decode_render_compare(x, w, bias, p, q, s, obs, pts, R, thr, cam, base=base,
                                                   depth_weight=1.0, pc_weight=3.0)
    up = torch.tensor([1.0, 0.5, 2.0], device=dev)
    (loss * up).sum().backward()

    x2, p2, q2, s2 = leaves()
    grids = decoder_tail(x2, w, bias, R, base)
    ld, depth2, n2 = render_and_compare(grids, p2, q2, (1.0 / s2).contiguous(), obs, thr, cam)
    ref = torch.nan_to_num(ld, nan=0.0)
    if with_points:
        ref = ref + 3.0 * losses.point_loss(pts, p2, q2, s2, grids)
    (ref * up).sum().backward()

    assert torch.equal(depth, depth2) and torch.equal(n, n2) and int(n.min()) > 0
    assert torch.allclose(loss, ref, rtol=1e-5, atol=1e-7) and torch.allclose(loss_d, ld, rtol=1e-5)
    for got, want in ((x.grad, x2.grad), (p.grad, p2.grad), (q.grad, q2.grad), (s.grad, s2.grad)):
        assert rel(got.cpu().numpy(), want.cpu().numpy().astype(np.float64)) < 1e-3
    # evaluation only (no gradient requested) takes the compare-forward path
    with torch.no_grad():
        l3, d3, _, _ = decode_render_compare(x.detach(), w, bias, p.detach(), q.detach(), s.detach(), obs,
                                             pts, R, thr, cam, base=base)
    assert torch.equal(d3, depth) and torch.allclose(l3, loss.detach(), rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_optimizer_uses_the_fused_chain_and_descends(cuda_device):
    from sdfest_b200 import synthetic as syn
    from sdfest_b200.differentiable_renderer import Camera, render_depth_batched
    from sdfest_b200.estimation import HypothesisOptimizer

    dev = cuda_device
    torch.manual_seed(0)
    B, R, W, H, thr = 4, 64, 160, 120, 0.005
    cam = Camera(W, H, 80.0, 80.0, 80.0, 60.0, pixel_center=0.5)
    hyp = syn.make_hypotheses(B, seed=5, device=dev)
    base = syn.sdf_mug(R, dev)
    obs = render_depth_batched(base, hyp["position"][:1], hyp["orientation"][:1], hyp["inv_scale"][:1],
                               thr, cam)[0].contiguous()
    dec = syn.residual_decoder(R, dev, base)
    opt = HypothesisOptimizer(cam, thr, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                              latent=torch.zeros(B, 8, device=dev), decoder=dec)
    first = opt.step().clone()
    for _ in range(25):
        last = opt.step()
    assert torch.isfinite(last).all() and float(last.mean()) < float(first.mean())
    assert float(opt.latent.abs().max()) > 0  # the latent moved: gradients reached the trunk
    opt2 = HypothesisOptimizer(cam, thr, obs, hyp["position"], hyp["orientation"], 1.0 / hyp["inv_scale"],
                               latent=torch.zeros(B, 8, device=dev), decoder=dec)
    opt2.capture()
    l2 = opt2.step()
    assert torch.isfinite(l2).all()
