"""numpy restatement of the reference decoder's last two operators -- TEST INFRASTRUCTURE ONLY.

    sdfest/vae/sdf_vae.py:235-244   interpolate(x, size=(R,R,R), mode="trilinear", align_corners=False)
    sdfest/vae/sdf_vae.py:245-247   Conv3d(C -> 1, kernel_size=1)(.)   (relu: false for the last stage)

The arithmetic lives in a third-party dependency of the reference, PyTorch (ATen
``upsample_trilinear3d`` / ``UpSampleTrilinear3d.cu`` and a 1x1x1 convolution; the container has
torch 2.11.0).  Its published algorithm for ``align_corners=False`` without an explicit
scale_factor is restated here: along each axis, output index o reads source indices
    src = max(0, (S / R) * (o + 0.5) - 0.5)        (area_pixel_compute_source_index, fp32)
    i0 = int(src),  i1 = i0 + (i0 < S - 1),  l1 = src - i0,  l0 = 1 - l1
and the value is the l-weighted sum over the 8 (i, j, k) corners; the convolution then contracts
the channels and adds the bias.  Sums are evaluated in float64 here (the oracle does not mirror
any particular fp32 summation order -- torch's, cuDNN's and the CUDA kernel's all differ).

Pinned by tests/test_oracle_golden.py against (1) vectors recorded from the reference's own
``SDFDecoder`` (tests/golden/decoder_tail_*.npz, made by tests/golden/make_golden_decoder.py: the
trained mug model of the reference's tests, and a random-init decoder with odd sizes) and (2)
``torch.nn.functional.interpolate`` + ``conv3d`` evaluated on the CPU in the test itself.
"""
from __future__ import annotations

import numpy as np


def axis_weights(S: int, R: int) -> np.ndarray:
    """(R, S) float64 matrix W with W[o, i0] += l0, W[o, i1] += l1 (indices / lambdas in fp32)."""
    ratio = np.float32(S) / np.float32(R)
    o = np.arange(R, dtype=np.float32)
    src = ratio * (o + np.float32(0.5)) - np.float32(0.5)
    src = np.maximum(src, np.float32(0.0)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), S - 1)
    i1 = i0 + (i0 < S - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    W = np.zeros((R, S), dtype=np.float64)
    np.add.at(W, (np.arange(R), i0), l0.astype(np.float64))
    np.add.at(W, (np.arange(R), i1), l1.astype(np.float64))
    return W


def tail_forward(x, weight, bias, R: int) -> np.ndarray:
    """x (B,C,S,S,S), weight (C,), bias scalar or None -> grids (B,R,R,R) float32."""
    x = np.asarray(x, dtype=np.float64)
    S = x.shape[-1]
    W = axis_weights(S, R)
    v = np.einsum("c,bcijk->bijk", np.asarray(weight, dtype=np.float64), x)
    out = np.einsum("xi,yj,zk,bijk->bxyz", W, W, W, v, optimize=True)
    if bias is not None:
        out = out + float(bias)
    return out.astype(np.float32)


def tail_backward(g, weight, S: int, coef=None, g_extra=None) -> np.ndarray:
    """Adjoint w.r.t. x: g (B,R,R,R) [scaled per hypothesis by coef (B,)] [+ g_extra] -> (B,C,S,S,S)."""
    g = np.asarray(g, dtype=np.float64)
    if coef is not None:
        g = g * np.asarray(coef, dtype=np.float64)[:, None, None, None]
    if g_extra is not None:
        g = g + np.asarray(g_extra, dtype=np.float64)
    R = g.shape[-1]
    W = axis_weights(S, R)
    gv = np.einsum("xi,yj,zk,bxyz->bijk", W, W, W, g, optimize=True)
    return np.einsum("c,bijk->bcijk", np.asarray(weight, dtype=np.float64), gv).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# trunk stages (sdfest/vae/sdf_vae.py:225-247): interpolate -> Conv3d(k=3, valid) -> optional ReLU
# ---------------------------------------------------------------------------------------------
def upsample(x, out_size: int) -> np.ndarray:
    """x (..., S, S, S) -> (..., U, U, U) float64: ATen trilinear, align_corners = False."""
    x = np.asarray(x, dtype=np.float64)
    W = axis_weights(x.shape[-1], out_size)
    return np.einsum("xi,yj,zk,...ijk->...xyz", W, W, W, x, optimize=True)


def upsample_backward(g, in_size: int) -> np.ndarray:
    g = np.asarray(g, dtype=np.float64)
    W = axis_weights(in_size, g.shape[-1])
    return np.einsum("xi,yj,zk,...xyz->...ijk", W, W, W, g, optimize=True)


def conv3d(x, weight, bias=None, relu=False) -> np.ndarray:
    """'valid' cross-correlation as torch.nn.Conv3d: x (B,Ci,U,U,U), weight (Co,Ci,k,k,k)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    k = w.shape[-1]
    O = x.shape[-1] - k + 1
    out = np.zeros((x.shape[0], w.shape[0], O, O, O))
    for dx in range(k):
        for dy in range(k):
            for dz in range(k):
                out += np.einsum("oc,bcxyz->boxyz", w[:, :, dx, dy, dz],
                                 x[:, :, dx:dx + O, dy:dy + O, dz:dz + O])
    if bias is not None:
        out += np.asarray(bias, dtype=np.float64)[None, :, None, None, None]
    return np.maximum(out, 0.0) if relu else out


def conv3d_backward_data(g, y, weight) -> np.ndarray:
    """Gradient w.r.t. the input of conv3d(+ReLU): g, y (B,Co,O,O,O); y = forward output (ReLU mask,
    None = no ReLU)."""
    g = np.asarray(g, dtype=np.float64)
    if y is not None:
        g = g * (np.asarray(y) > 0)
    w = np.asarray(weight, dtype=np.float64)
    k = w.shape[-1]
    O = g.shape[-1]
    U = O + k - 1
    gx = np.zeros((g.shape[0], w.shape[1], U, U, U))
    for dx in range(k):
        for dy in range(k):
            for dz in range(k):
                gx[:, :, dx:dx + O, dy:dy + O, dz:dz + O] += np.einsum("oc,boxyz->bcxyz", w[:, :, dx, dy, dz], g)
    return gx


def decoder_forward(z, fc, conv, volume_size: int) -> np.ndarray:
    """The whole reference decoder (sdf_vae.py:217-259) in float64.  fc: [(W (out,in), b)], conv:
    [(in_size, W (Co,Ci,k,k,k), b, relu)]."""
    out = np.asarray(z, dtype=np.float64)
    for W, b in fc:
        out = np.maximum(out @ np.asarray(W, dtype=np.float64).T + np.asarray(b, dtype=np.float64), 0.0)
    s0, c0 = conv[0][0], conv[0][1].shape[1]
    out = out.reshape(-1, c0, s0, s0, s0)
    for in_size, W, b, relu in conv:
        if out.shape[2] != in_size:
            out = upsample(out, in_size)
        if W.shape[-1] == 1:
            out = np.einsum("oc,bcxyz->boxyz", np.asarray(W, dtype=np.float64)[:, :, 0, 0, 0], out) \
                + np.asarray(b, dtype=np.float64)[None, :, None, None, None]
            out = np.maximum(out, 0.0) if relu else out
        else:
            out = conv3d(out, W, b, relu)
    if out.shape[2] != volume_size:
        out = upsample(out, volume_size)
    return out
