"""BASELINE config 4: a hypothesis sweep -- N_TOTAL pose/scale hypotheses on synthetic mug / bowl /
bottle SDFs, sharded over the GPUs of one box (strong scaling: the total is fixed), ITER Adam steps
each, per-iteration all_gather of the per-hypothesis losses over NCCL, global arg-min at the end.

    python scripts/gpu_sweep.py [tag]                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \\
        --master-port 29533 scripts/gpu_sweep.py [tag]                  # G GPUs

Every rank owns a contiguous shard (estimation.shard_range); within the shard the hypotheses of one
category share ONE grid (sdf_stride = 0), so a rank runs three fused optimisers (3 launches per
iteration each), the rank's whole iteration replayed as one CUDA graph.  Prints one JSON line on rank 0 and writes
gpurun_out/<tag>_sweep_n<G>.json.  Time = CUDA events around the ITER iterations, max over ranks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfest_b200 import synthetic as syn  # noqa: E402
from sdfest_b200.differentiable_renderer import Camera, render_depth_batched  # noqa: E402
from sdfest_b200.estimation import HypothesisOptimizer  # noqa: E402
from sdfest_b200.estimation.hypotheses import gather_losses, shard_range  # noqa: E402

N_TOTAL = int(os.environ.get("SWEEP_N", "4096"))
ITER = int(os.environ.get("SWEEP_ITER", "20"))
W, H, R, THR = 640, 480, 64, 0.005

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

cam = Camera(W, H, 320.0, 320.0, 320.0, 240.0, pixel_center=0.5)
hyp = syn.make_hypotheses(N_TOTAL, seed=0, device=dev)  # same set on every rank, sharded below
cat_of = torch.arange(N_TOTAL, device=dev) % 3          # mug / bowl / bottle, a third each
base = syn.make_hypotheses(1, seed=0, device=dev)
obs = render_depth_batched(syn.category_grid("mug", R, dev)[None], base["position"], base["orientation"],
                           base["inv_scale"], THR, cam)[0].contiguous()
lo, hi = shard_range(N_TOTAL, rank, world)
opts, index = [], []
for c, name in enumerate(syn.CATEGORIES):
    sel = torch.nonzero(cat_of[lo:hi] == c).flatten() + lo
    if sel.numel() == 0:
        continue
    grid = syn.category_grid(name, R, dev)[None].contiguous()
    opt = HypothesisOptimizer(cam, THR, obs, hyp["position"][sel], hyp["orientation"][sel],
                              1.0 / hyp["inv_scale"][sel], sdf=grid, max_points=20000)
    opts.append(opt)
    index.append(sel)
index = torch.cat(index)
local_losses = torch.empty(hi - lo, device=dev)


# The categories are independent optimisers: each iterates on its own stream (forked from / joined to the
# calling stream), so inside the captured graph they are parallel branches and one category's tail -- the last
# CTAs of its render launch, its small step kernel -- overlaps the next category's work.  SWEEP_SERIAL=1 keeps
# them on one stream (the A/B: profiles/*_sweep_branches.json).
SERIAL = os.environ.get("SWEEP_SERIAL", "0") == "1"
branches = [torch.cuda.Stream(dev) for _ in opts]


def local_iteration():
    off = 0
    main = torch.cuda.current_stream()
    for opt, br in zip(opts, branches):
        n = opt.position.shape[0]
        if SERIAL:
            local_losses[off:off + n] = opt.step()
        else:
            br.wait_stream(main)
            with torch.cuda.stream(br):
                local_losses[off:off + n] = opt.step()
        off += n
    if not SERIAL:
        for br in branches:
            main.wait_stream(br)


# the rank's whole iteration (three categories x three launches + the loss concatenation) as ONE graph
side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        local_iteration()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    local_iteration()


SIZES = [h - l for l, h in (shard_range(N_TOTAL, r, world) for r in range(world))]


def iteration():
    graph.replay()
    return gather_losses(local_losses, sizes=SIZES)  # the path's only exchange: (hi - lo) floats per rank


for _ in range(3):
    iteration()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(ITER):
    allv = iteration()
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], device=dev)
all_index = index
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sizes = [shard_range(N_TOTAL, r, world) for r in range(world)]
    if len({h - l for l, h in sizes}) == 1:
        out = torch.empty(world * index.numel(), dtype=index.dtype, device=dev)
        dist.all_gather_into_tensor(out, index)
        all_index = out
best = int(torch.argmin(torch.nan_to_num(allv, nan=float("inf"))))
line = {"workload": f"C4 sweep: {N_TOTAL} pose/scale hypotheses on mug/bowl/bottle 64^3 grids, 640x480, "
                    f"{ITER} fused Adam iterations, all_gather of losses every iteration",
        "n_gpus": world, "hypotheses_total": N_TOTAL, "hypotheses_per_gpu": hi - lo, "iterations": ITER,
        "ms_per_iteration": float(ms) / ITER, "hyp_iter_per_s": N_TOTAL * ITER / (float(ms) * 1e-3),
        "scaling": "strong", "category_branches": "serial" if SERIAL else "parallel", "best_hypothesis": int(all_index[best]) if best < all_index.numel() else best,
        "best_loss": float(allv[best]), "mean_loss": float(torch.nan_to_num(allv).mean())}
if rank == 0:
    tag = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", f"{tag}_sweep_n{world}.json"), "w"), indent=1)
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
