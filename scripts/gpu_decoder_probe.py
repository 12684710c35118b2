"""Where does the decoder's time go?  Per-operator CUDA time of SDFDecoder forward+backward at
B hypotheses (torch profiler), to decide what to fuse.  Writes gpurun_out/<tag>_decoder_ops.txt."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200.estimation import SDFDecoder  # noqa: E402

B = int(os.environ.get("LOOP_B", "64"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = SDFDecoder(64).to(dev).eval()
for p in dec.parameters():
    p.requires_grad_(False)
lat = torch.randn(B, 8, device=dev, requires_grad=True)
g = torch.randn(B, 1, 64, 64, 64, device=dev)


def step():
    out = dec(lat)
    out.backward(g)
    lat.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90)
tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"{tag}_decoder_ops.txt"), "w").write(txt)
print(txt)
for cudnn in (True, False):
    torch.backends.cudnn.enabled = cudnn
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step()
    b.record()
    torch.cuda.synchronize()
    print(f"cudnn={cudnn}: decoder fwd+bwd {a.elapsed_time(b) / 10:.3f} ms at B={B}")
