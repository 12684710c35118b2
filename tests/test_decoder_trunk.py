"""Decoder trunk stages (reference sdfest/vae/sdf_vae.py:217-259): trilinear resize + 3x3x3
convolution + ReLU as CUDA kernels, and the whole decoder assembled from them.

CPU: the numpy oracle (oracle/decoder_tail.py: upsample, conv3d, decoder_forward) against a whole
reference SDFDecoder recorded in tests/golden/decoder_full_small.npz and against torch's CPU
operators.  GPU: kernels (C ABI) against oracle, golden and torch CUDA (TF32 off).
Floating point, reassociated sums: 1e-5 of the largest magnitude forward, 1e-4 backward."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import decoder_tail as odt
from sdfest_b200 import _lib
from sdfest_b200.estimation import FusedTailDecoder, SDFDecoder, trunk_stage

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD_TOL, BWD_TOL = 1e-5, 1e-4


def rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def golden_decoder():
    z = np.load(os.path.join(GOLDEN, "decoder_full_small.npz"))
    fc = [(z[f"fc{i}_w"], z[f"fc{i}_b"]) for i in range(2)]
    conv = [(int(z["in_sizes"][i]), z[f"conv{i}_w"], z[f"conv{i}_b"], bool(z["relus"][i])) for i in range(4)]
    return z, fc, conv


def module_from_golden(z):
    """This package's SDFDecoder carrying the golden reference decoder's weights."""
    conv_info = tuple((int(z["in_sizes"][i]), z[f"conv{i}_w"].shape[1], z[f"conv{i}_w"].shape[0],
                       z[f"conv{i}_w"].shape[-1], bool(z["relus"][i])) for i in range(4))
    dec = SDFDecoder(20, latent_size=3, fc=(10, 8 * 4 ** 3), conv=conv_info)
    with torch.no_grad():
        for i, layer in enumerate(dec.fc):
            layer.weight.copy_(torch.tensor(z[f"fc{i}_w"]))
            layer.bias.copy_(torch.tensor(z[f"fc{i}_b"]))
        for i, layer in enumerate(dec.conv):
            layer.weight.copy_(torch.tensor(z[f"conv{i}_w"]))
            layer.bias.copy_(torch.tensor(z[f"conv{i}_b"]))
    return dec.eval()


def test_oracle_reproduces_the_reference_decoder():
    z, fc, conv = golden_decoder()
    assert rel(odt.decoder_forward(z["z"], fc, conv, 20)[:, 0], z["out"]) < 1e-6
    # this package's torch SDFDecoder is the same function of the same weights
    dec = module_from_golden(z)
    zt = torch.tensor(z["z"], requires_grad=True)
    out = dec(zt)
    assert rel(out.detach().numpy()[:, 0], z["out"]) < 1e-6
    out.backward(torch.tensor(z["g"])[:, None])
    assert rel(zt.grad.numpy(), z["g_z"]) < 1e-5


@pytest.mark.parametrize("Ci,Co,S,U,relu", [(8, 4, 14, 32, True), (16, 8, 6, 16, True), (16, 16, 8, 8, False),
                                            (4, 32, 5, 9, True)])
def test_oracle_stage_matches_torch_cpu(Ci, Co, S, U, relu):
    rng = np.random.default_rng(Ci * Co + S)
    x = torch.tensor(rng.standard_normal((2, Ci, S, S, S)), dtype=torch.float64, requires_grad=True)
    w = torch.tensor(rng.standard_normal((Co, Ci, 3, 3, 3)) * 0.1, dtype=torch.float64)
    b = torch.tensor(rng.standard_normal(Co) * 0.1, dtype=torch.float64)
    u = F.interpolate(x, size=(U,) * 3, mode="trilinear", align_corners=False) if S != U else x
    y = F.conv3d(u, w, b)
    y = torch.relu(y) if relu else y
    g = torch.tensor(rng.standard_normal(tuple(y.shape)), dtype=torch.float64)
    y.backward(g)
    xu = odt.upsample(x.detach().numpy(), U) if S != U else x.detach().numpy()
    yo = odt.conv3d(xu, w.numpy(), b.numpy(), relu)
    assert rel(yo, y.detach().numpy()) < 1e-6  # interpolation lambdas are fp32 in ATen and here
    gu = odt.conv3d_backward_data(g.numpy(), yo if relu else None, w.numpy())
    gx = odt.upsample_backward(gu, S) if S != U else gu
    assert rel(gx, x.grad.numpy()) < 1e-6


def test_trunk_selection_and_argument_errors():
    dec = SDFDecoder(64)
    assert FusedTailDecoder(dec).trunk_impl == "cuda"
    assert FusedTailDecoder(SDFDecoder(64), trunk="torch").trunk_impl == "torch"
    odd = ((8, 16, 6, 3, True), (16, 6, 4, 3, True), (30, 4, 1, 1, False))  # 6 channels: not supported
    f = FusedTailDecoder(SDFDecoder(30, conv=odd))
    assert f.trunk_impl == "torch"
    with pytest.raises(ValueError, match="trunk='cuda'"):
        FusedTailDecoder(SDFDecoder(30, conv=odd), trunk="cuda")
    # on the CPU the wrapper's trunk is plain torch and equals the wrapped decoder's
    z = torch.randn(2, 8)
    fused = FusedTailDecoder(dec)
    with torch.no_grad():
        x = fused.trunk(z)
        w, b = fused.tail_parameters()
        y = F.conv3d(F.interpolate(x, size=(64,) * 3, mode="trilinear", align_corners=False),
                     w.view(1, -1, 1, 1, 1), b)
        assert torch.allclose(y, dec(z), atol=1e-6)
    lib = _lib.lib()
    assert lib.sdfr_conv3d_forward(None, 0, 8, 16, None, None, 4, 3, 1, None, None) == 0
    assert lib.sdfr_conv3d_forward(None, 1, 8, 16, None, None, 4, 3, 1, None, None) == -1
    assert lib.sdfr_conv3d_forward(None, 1, 8, 16, None, None, 5, 3, 1, None, None) == -2
    assert lib.sdfr_conv3d_forward(None, 1, 8, 16, None, None, 4, 5, 1, None, None) == -2
    assert lib.sdfr_conv3d_forward(None, 1, 8, 2, None, None, 4, 3, 1, None, None) == -2
    assert lib.sdfr_conv3d_backward_data(None, None, 1, 6, 16, None, 4, 3, None, None) == -2
    assert lib.sdfr_conv3d_backward_data(None, None, 1, 8, 16, None, 4, 3, None, None) == -1
    assert lib.sdfr_upsample3d_forward(None, 0, 6, 16, None, None) == 0
    assert lib.sdfr_upsample3d_forward(None, 3, 6, 16, None, None) == -1
    assert lib.sdfr_upsample3d_backward(None, 3, 6, 300, None, None) == -2


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("Ci,Co,S,U,relu,B", [(8, 4, 14, 32, True, 3), (16, 8, 6, 16, True, 2),
                                              (16, 16, 8, 8, True, 2), (32, 16, 4, 4, False, 2),
                                              (4, 32, 5, 9, True, 1), (8, 8, 20, 37, False, 1),
                                              (16, 4, 3, 3, True, 1)])
def test_stage_kernels_match_oracle_and_torch(cuda_device, Ci, Co, S, U, relu, B):
    rng = np.random.default_rng(Ci * Co + S + U)
    dev = cuda_device
    x = torch.tensor(rng.standard_normal((B, Ci, S, S, S)), dtype=torch.float32, device=dev, requires_grad=True)
    w = torch.tensor(rng.standard_normal((Co, Ci, 3, 3, 3)) * 0.1, dtype=torch.float32, device=dev)
    b = torch.tensor(rng.standard_normal(Co) * 0.1, dtype=torch.float32, device=dev)
    y = trunk_stage(x, w, b, U, relu)
    O = U - 2
    assert tuple(y.shape) == (B, Co, O, O, O)
    g = torch.tensor(rng.standard_normal(tuple(y.shape)), dtype=torch.float32, device=dev)
    y.backward(g)
    xn, wn, bn = x.detach().cpu().numpy(), w.cpu().numpy(), b.cpu().numpy()
    xu = odt.upsample(xn, U) if S != U else xn.astype(np.float64)
    yo = odt.conv3d(xu, wn, bn, relu)
    got_y = y.detach().cpu().numpy()
    assert rel(got_y, yo) < FWD_TOL
    gu = odt.conv3d_backward_data(g.cpu().numpy(), got_y if relu else None, wn)
    gx = odt.upsample_backward(gu, S) if S != U else gu
    assert rel(x.grad.cpu().numpy(), gx) < BWD_TOL
    # torch's CUDA operators (TF32 off)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x2 = x.detach().clone().requires_grad_(True)
        u2 = F.interpolate(x2, size=(U,) * 3, mode="trilinear", align_corners=False) if S != U else x2
        y2 = F.conv3d(u2, w, b)
        y2 = torch.relu(y2) if relu else y2
        y2.backward(g)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert rel(got_y, y2.detach().cpu().numpy().astype(np.float64)) < FWD_TOL
    # ReLU masks can differ where |pre-activation| ~ 1e-7: compare where both agree on the sign
    assert rel(x.grad.cpu().numpy(), x2.grad.cpu().numpy().astype(np.float64)) < 1e-3


@pytest.mark.gpu
def test_whole_decoder_on_cuda_kernels_matches_reference_golden(cuda_device):
    z, fc, conv = golden_decoder()
    dec = module_from_golden(z).to(cuda_device)
    fused = FusedTailDecoder(dec, trunk="cuda")
    assert fused.trunk_impl == "cuda"
    zt = torch.tensor(z["z"], device=cuda_device, requires_grad=True)
    out = fused(zt)
    assert tuple(out.shape) == (2, 1, 20, 20, 20)
    assert rel(out.detach().cpu().numpy()[:, 0], z["out"]) < FWD_TOL
    out.backward(torch.tensor(z["g"], device=cuda_device)[:, None])
    assert rel(zt.grad.cpu().numpy(), z["g_z"]) < 1e-3


@pytest.mark.gpu
def test_cuda_trunk_equals_torch_trunk_on_the_mug_architecture(cuda_device):
    torch.manual_seed(4)
    dec = SDFDecoder(64).to(cuda_device).eval()
    with torch.no_grad():
        for layer in dec.conv:
            layer.bias.add_(0.05)
    a = FusedTailDecoder(dec, trunk="cuda")
    b = FusedTailDecoder(dec, trunk="torch", channels_last=False)
    z1 = torch.randn(5, 8, device=cuda_device, requires_grad=True)
    z2 = z1.detach().clone().requires_grad_(True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xa, xb = a.trunk(z1), b.trunk(z2)
        g = torch.randn_like(xa)
        xa.backward(g)
        xb.backward(g)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert tuple(xa.shape) == (5, 4, 30, 30, 30)
    assert rel(xa.detach().cpu().numpy(), xb.detach().cpu().numpy().astype(np.float64)) < FWD_TOL
    assert rel(z1.grad.cpu().numpy(), z2.grad.cpu().numpy().astype(np.float64)) < 1e-3
