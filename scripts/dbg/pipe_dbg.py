import importlib, os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_loader, test_dropin_reference as t
from oracle import build_ref
from sdfest_b200.estimation import SDFPipeline
from sdfest_b200.differentiable_renderer import set_empty_space_policy
dev = torch.device("cuda:0")
ext = build_ref.load_module()
vae_path, vae_yaml = os.path.join(ref_loader.FIXTURES, "mug.pt"), os.path.join(ref_loader.FIXTURES, "mug.yaml")
init_path = os.path.join(tempfile.mkdtemp(), "init.pt")
depth, q_true = t._observation(dev, vae_path, vae_yaml)
cfg = t._pipeline_config(init_path, vae_yaml, vae_path, 100)
ref_loader.load_reference(ext)
setup = importlib.import_module("sdfest.estimation.simple_setup")
t._make_init_weights(setup, cfg, q_true, init_path)
pipe = setup.SDFPipeline(cfg)
vae, init_network = pipe.vae, pipe.init_network
ref_loader.purge()
torch.backends.cudnn.enabled = True
for policy in ("auto", "off", "auto"):
    set_empty_space_policy(policy)
    for prof in (False, True, False):
        mine = SDFPipeline(dict(cfg, relative_inlier_threshold=0.03, profile=prof), vae, init_network)
        ts = []
        for _ in range(4):
            d = depth.clone(); torch.cuda.synchronize(); t0 = time.perf_counter()
            mine(d, d > 0, None); torch.cuda.synchronize(); ts.append(round((time.perf_counter() - t0) * 1e3, 1))
        print(policy, "profile", prof, ts, mine.last_timings and {k: round(v, 1) for k, v in mine.last_timings.items()}, flush=True)
