"""Decoder trunk timing under different cuDNN settings (benchmark mode, TF32, channels_last_3d)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200.estimation import FusedTailDecoder, SDFDecoder  # noqa: E402

B = int(os.environ.get("LOOP_B", "64"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = FusedTailDecoder(SDFDecoder(64)).to(dev).eval()
lat = torch.randn(B, 8, device=dev, requires_grad=True)
g = torch.randn(B, 1, 64, 64, 64, device=dev)


def step():
    out = dec(lat)
    out.backward(g)
    lat.grad = None


def run(label):
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        step()
    b.record()
    torch.cuda.synchronize()
    print(f"{label}: fused-tail decoder fwd+bwd {a.elapsed_time(b) / 20:.3f} ms at B={B}", flush=True)


run("default")
torch.backends.cudnn.benchmark = True
run("cudnn.benchmark")
torch.backends.cudnn.allow_tf32 = False
run("cudnn.benchmark, no tf32")
torch.backends.cudnn.allow_tf32 = True
dec = dec.to(memory_format=torch.channels_last_3d)
run("cudnn.benchmark + channels_last_3d weights")
