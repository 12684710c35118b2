"""Golden vectors for the decoder tail, generated from the REFERENCE's own decoder class.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden_decoder.py

Imports ``sdfest.vae.sdf_vae`` from /root/reference (imports cleanly: torch + yaml only) and
records, for the last two operators of ``SDFDecoder.forward`` (sdf_vae.py:235-247:
``interpolate(..., mode="trilinear", align_corners=False)`` then ``Conv3d(C -> 1, kernel_size=1)``),
the tensor that enters them, the grid that leaves them, and -- through the reference's own
autograd graph -- the gradient w.r.t. that input for a seeded upstream gradient.

 * decoder_tail_mug_z0.npz : the trained mug VAE shipped with the reference's tests
   (tests/initilization/vae_model/mug.pt) decoded at z = 0: x (4,30,30,30), weight (4,), bias ();
   the output is tests/golden/mug_z0_sdf.npz (already a fixture of the renderer tests), plus
   g_x = d<g, sdf>/dx for g = seeded N(0,1), stored for every 5th x-plane only (keeps it small).
 * decoder_tail_small.npz  : a random-init reference SDFDecoder with odd sizes (C=3, S=5 -> R=12,
   batch 2): x, weight, bias, out, g, g_x in full.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
from sdfest.vae.sdf_vae import SDFVAE, SDFDecoder  # noqa: E402


def tail_io(decoder, z, g):
    """Run decoder(z) with a hook on the input of the last conv stage's interpolation."""
    captured = {}
    n_stage = len(decoder._conv_layers)
    orig = torch.nn.functional.interpolate

    def spy(inp, *a, **kw):
        captured["calls"] = captured.get("calls", 0) + 1
        captured["last_in"] = inp
        if inp.requires_grad:
            inp.retain_grad()
        return orig(inp, *a, **kw)

    torch.nn.functional.interpolate = spy
    try:
        out = decoder(z)
    finally:
        torch.nn.functional.interpolate = orig
    x = captured["last_in"]
    out.backward(g)
    return x, out, x.grad, n_stage


def mug():
    cfg = yaml.safe_load(open(os.path.join(REF, "tests/initilization/vae_model/mug.yaml")))
    vae = SDFVAE(sdf_size=64, latent_size=cfg["latent_size"], encoder_dict=cfg["encoder"],
                 decoder_dict=cfg["decoder"], device="cpu", tsdf=cfg["tsdf"])
    vae.load_state_dict(torch.load(os.path.join(REF, "tests/initilization/vae_model/mug.pt"),
                                   map_location="cpu"))
    vae.eval()
    dec = vae.decoder
    z = torch.zeros(1, cfg["latent_size"], requires_grad=True)
    g = torch.as_tensor(np.random.default_rng(21).standard_normal((1, 1, 64, 64, 64)), dtype=torch.float32)
    x, out, gx, _ = tail_io(dec, z, g)
    ref = np.load(os.path.join(HERE, "mug_z0_sdf.npz"))["sdf"]
    assert np.array_equal(out.detach().numpy()[0, 0], ref), "mug_z0_sdf.npz is not this decode"
    conv = dec._conv_layers[-1]
    planes = np.arange(0, x.shape[2], 5)
    path = os.path.join(HERE, "decoder_tail_mug_z0.npz")
    np.savez_compressed(path, x=x.detach().numpy()[0], weight=conv.weight.detach().numpy().reshape(-1),
                        bias=conv.bias.detach().numpy().reshape(()), g_seed=np.int64(21),
                        g_x_planes=planes, g_x=gx.numpy()[0][:, planes])
    print("decoder_tail_mug_z0:", tuple(x.shape), "->", tuple(out.shape), f"{os.path.getsize(path) / 1024:.0f} KiB")


def small():
    torch.manual_seed(5)
    dec = SDFDecoder(volume_size=12, latent_size=4, fc_layers=[{"out": 2 * 4 ** 3}],
                     conv_layers=[{"in_size": 4, "in_channels": 2, "out_channels": 3, "kernel_size": 1, "relu": True},
                                  {"in_size": 7, "in_channels": 3, "out_channels": 3, "kernel_size": 3, "relu": True},
                                  {"in_size": 12, "in_channels": 3, "out_channels": 1, "kernel_size": 1, "relu": False}])
    dec.eval()
    z = torch.randn(2, 4, requires_grad=True)
    g = torch.randn(2, 1, 12, 12, 12)
    x, out, gx, _ = tail_io(dec, z, g)
    conv = dec._conv_layers[-1]
    path = os.path.join(HERE, "decoder_tail_small.npz")
    np.savez_compressed(path, x=x.detach().numpy(), weight=conv.weight.detach().numpy().reshape(-1),
                        bias=conv.bias.detach().numpy().reshape(()), out=out.detach().numpy()[:, 0],
                        g=g.numpy()[:, 0], g_x=gx.numpy())
    print("decoder_tail_small:", tuple(x.shape), "->", tuple(out.shape), f"{os.path.getsize(path) / 1024:.0f} KiB")


def full_decoder():
    """A whole (small, random-init) reference SDFDecoder with the shipped models' structure:
    fc stack, three 3x3x3 stages with interpolation in between, 1x1x1 tail; weights, latents, the
    decoded grids and d<g, out>/dz recorded through the reference's own forward/autograd."""
    torch.manual_seed(9)
    conv_layers = [{"in_size": 4, "in_channels": 8, "out_channels": 16, "kernel_size": 3, "relu": True},
                   {"in_size": 7, "in_channels": 16, "out_channels": 8, "kernel_size": 3, "relu": True},
                   {"in_size": 11, "in_channels": 8, "out_channels": 4, "kernel_size": 3, "relu": True},
                   {"in_size": 20, "in_channels": 4, "out_channels": 1, "kernel_size": 1, "relu": False}]
    dec = SDFDecoder(volume_size=20, latent_size=3, fc_layers=[{"out": 10}, {"out": 8 * 4 ** 3}],
                     conv_layers=conv_layers)
    dec.eval()
    with torch.no_grad():  # default init gives mostly-dead ReLUs; spread the activations
        for layer in dec._conv_layers:
            layer.bias.add_(0.05)
    z = torch.randn(2, 3, requires_grad=True)
    g = torch.randn(2, 1, 20, 20, 20)
    out = dec(z)
    out.backward(g)
    blobs = {f"fc{i}_w": l.weight.detach().numpy() for i, l in enumerate(dec._fc_layers)}
    blobs.update({f"fc{i}_b": l.bias.detach().numpy() for i, l in enumerate(dec._fc_layers)})
    blobs.update({f"conv{i}_w": l.weight.detach().numpy() for i, l in enumerate(dec._conv_layers)})
    blobs.update({f"conv{i}_b": l.bias.detach().numpy() for i, l in enumerate(dec._conv_layers)})
    path = os.path.join(HERE, "decoder_full_small.npz")
    np.savez_compressed(path, z=z.detach().numpy(), g=g.numpy()[:, 0], out=out.detach().numpy()[:, 0],
                        g_z=z.grad.numpy(), in_sizes=np.array([c["in_size"] for c in conv_layers]),
                        relus=np.array([c["relu"] for c in conv_layers]), **blobs)
    print("decoder_full_small:", tuple(out.shape), f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    mug()
    small()
    full_decoder()
