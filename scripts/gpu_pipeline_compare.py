"""The reference's SDFPipeline.__call__ (estimation/simple_setup.py:213-600: initialisation network, then
max_iterations x {decode, render, losses, backward, Adam}) on ITS OWN CUDA extension, against this
package's SDFPipeline (fused iteration replayed from a CUDA graph) on the same trained mug VAE, the same
initialisation network and the same observation: wall time per call and per iteration, one hypothesis --
the unit the reference works in.  Needs baseline/_ref (oracle/install_reference.py) and oracle/_ref.

    python scripts/gpu_pipeline_compare.py [tag]    -> gpurun_out/<tag>_pipeline.json
"""
import importlib
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_loader  # noqa: E402
import test_dropin_reference as t  # noqa: E402

dev = torch.device("cuda:0")
out = {}
if not ref_loader.available():
    out["unavailable"] = "baseline/_ref not installed"
else:
    from oracle import build_ref
    from sdfest_b200.estimation import SDFPipeline

    ext = build_ref.load_module()
    vae_path, vae_yaml = os.path.join(ref_loader.FIXTURES, "mug.pt"), os.path.join(ref_loader.FIXTURES, "mug.yaml")
    init_path = os.path.join(tempfile.mkdtemp(), "init.pt")
    depth, q_true = t._observation(dev, vae_path, vae_yaml)
    cudnn = torch.backends.cudnn.enabled
    for iterations in (50, 100):
        cfg = t._pipeline_config(init_path, vae_yaml, vae_path, iterations)
        ref_loader.load_reference(ext)
        setup = importlib.import_module("sdfest.estimation.simple_setup")
        t._make_init_weights(setup, cfg, q_true, init_path)
        pipe = setup.SDFPipeline(cfg)

        def ref_call():
            d = depth.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = pipe(d, d > 0, torch.zeros(*d.shape, 3, device=dev))
            torch.cuda.synchronize()
            return time.perf_counter() - t0, res

        ref_call()
        ref_s = min(ref_call()[0] for _ in range(3))
        vae, init_network = pipe.vae, pipe.init_network
        ref_loader.purge()
        torch.backends.cudnn.enabled = cudnn  # the reference pipeline switches cuDNN off globally (:46)

        mine = SDFPipeline(dict(cfg, relative_inlier_threshold=0.03), vae, init_network)

        def my_call():
            d = depth.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = mine(d, d > 0, None)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, res

        first_s = my_call()[0]  # builds the optimiser, three eager iterations, graph capture
        my_s = min(my_call()[0] for _ in range(3))  # later calls replay the cached graph (reuse_graph)
        nocache = SDFPipeline(dict(cfg, relative_inlier_threshold=0.03, reuse_graph=False), vae, init_network)

        def nocache_call():
            d = depth.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nocache(d, d > 0, None)
            torch.cuda.synchronize()
            return time.perf_counter() - t0

        nocache_call()
        nocache_s = min(nocache_call() for _ in range(3))
        prof = SDFPipeline(dict(cfg, relative_inlier_threshold=0.03, profile=True), vae, init_network)
        d = depth.clone()
        prof(d, d > 0, None)
        d = depth.clone()
        prof(d, d > 0, None)
        out[f"iterations_{iterations}"] = {
            "reference_pipeline_on_its_extension_ms": ref_s * 1e3, "this_package_ms": my_s * 1e3,
            "this_package_first_call_ms": first_s * 1e3, "this_package_capture_every_call_ms": nocache_s * 1e3,
            "graph_reused": mine.last_optimizer.point_capacity > 0,
            "speedup": ref_s / my_s, "optimizer": mine.last_optimizer.optimizer_impl,
            "graph": mine.last_optimizer._graph is not None, "phases_ms": prof.last_timings}
    # two views of the same observation (second camera 3 cm to the side) and a point constraint, 50 iterations:
    # the reference's loop over views (:420-446) against the fused view loop; later calls replay the cached graph
    cfg = t._pipeline_config(init_path, vae_yaml, vae_path, 50)
    views = torch.stack([depth, depth])
    cam_p = torch.tensor([[0.0, 0.0, 0.0], [0.03, 0.0, 0.0]], device=dev)
    cam_q = torch.tensor([[0.0, 0.0, 0.0, 1.0], [0.0, 0.0, 0.0, 1.0]], device=dev)
    constraint = (torch.tensor([0.0, 0.0, 1.0]), torch.tensor([0.0, 0.0, 1.0]), 0.05)
    ref_loader.load_reference(ext)
    setup = importlib.import_module("sdfest.estimation.simple_setup")
    t._make_init_weights(setup, cfg, q_true, init_path)
    pipe = setup.SDFPipeline(cfg)

    def timed_call(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def ref_views():
        d = views.clone()
        pipe(d, d > 0, torch.zeros(*d.shape, 3, device=dev), camera_positions=cam_p, camera_orientations=cam_q,
             point_constraint=tuple(x.to(dev) if torch.is_tensor(x) else x for x in constraint))

    try:
        timed_call(ref_views)
        ref_views_s = min(timed_call(ref_views) for _ in range(3))
    except Exception as e:  # noqa: BLE001
        ref_views_s = None
        out["views_reference_error"] = str(e)[:300]
    vae, init_network = pipe.vae, pipe.init_network
    ref_loader.purge()
    torch.backends.cudnn.enabled = cudnn
    res = {}
    for name, extra in (("graph_reused", {}), ("capture_every_call", {"reuse_graph": False})):
        mine = SDFPipeline(dict(cfg, relative_inlier_threshold=0.03, **extra), vae, init_network)

        def my_views():
            d = views.clone()
            mine(d, d > 0, None, camera_positions=cam_p, camera_orientations=cam_q, point_constraint=constraint)

        first = timed_call(my_views)
        res[name] = {"first_call_ms": first * 1e3, "later_calls_ms": min(timed_call(my_views) for _ in range(3)) * 1e3}
    out["two_views_with_constraint_50_iterations"] = {
        "reference_pipeline_on_its_extension_ms": None if ref_views_s is None else ref_views_s * 1e3,
        "this_package": res,
        "speedup": None if ref_views_s is None else ref_views_s * 1e3 / res["graph_reused"]["later_calls_ms"]}
    a, b = out["iterations_50"], out["iterations_100"]
    out["per_iteration_ms"] = {
        "reference": (b["reference_pipeline_on_its_extension_ms"] - a["reference_pipeline_on_its_extension_ms"]) / 50,
        "this_package": (b["this_package_ms"] - a["this_package_ms"]) / 50}
    out["per_iteration_ms"]["speedup"] = out["per_iteration_ms"]["reference"] / out["per_iteration_ms"]["this_package"]
    out["what"] = ("one 640x480 observation of the reference's trained mug VAE, one hypothesis; wall clock around the "
                   "pipeline call, best of 3; per-iteration = difference of the 100- and 50-iteration calls / 50")
tag = sys.argv[1] if len(sys.argv) > 1 else "cmp"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_pipeline.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
