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
