"""Golden vectors of the reference's quaternion helpers for the view loop.

    python tests/golden/make_golden_views.py        (build container: /root/reference mounted)

``sdfest/initialization/quaternion_utils.py`` is pure torch and loads by file path.  Stored in
``views.npz``: seeded unit quaternions / points and the outputs of quaternion_multiply,
quaternion_apply (unit and non-unit quaternions), quaternion_invert, ``losses.point_constraint_loss``
(estimation/losses.py:138-153) with its gradient, plus the camera-frame poses of the reference's view loop
(estimation/simple_setup.py:423-431, the three lines evaluated with those helpers).
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location(
    "ref_quaternion_utils", "/root/reference/sdfest/initialization/quaternion_utils.py")
qu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(qu)

g = torch.Generator().manual_seed(9)
unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, 4, generator=g, dtype=torch.float64), dim=-1)  # noqa: E731
q1, q2, pts = unit(7), unit(7), torch.randn(7, 3, generator=g, dtype=torch.float64)
V, B = 3, 5
cam_p, cam_q = torch.randn(V, 3, generator=g, dtype=torch.float64), unit(V)
pos, ori = torch.randn(B, 3, generator=g, dtype=torch.float64), unit(B)
pos_c, ori_c = [], []
for v in range(V):  # simple_setup.py:423-431
    q_w2c = qu.quaternion_invert(cam_q[v])
    pos_c.append(qu.quaternion_apply(q_w2c, pos - cam_p[v]))
    ori_c.append(qu.quaternion_multiply(q_w2c, ori))
# non-unit quaternions (the reference does not normalise) and the point-constraint loss of
# estimation/losses.py:138-153, its source executed unchanged with the helpers above injected
import ast  # noqa: E402
import textwrap  # noqa: E402

src = open("/root/reference/sdfest/estimation/losses.py").read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "point_constraint_loss")
ns = {"torch": torch, "quaternion_utils": qu}
exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
q_raw = torch.randn(6, 4, generator=g, dtype=torch.float64) * 1.3
source, target = torch.randn(3, generator=g, dtype=torch.float64), torch.randn(3, generator=g, dtype=torch.float64)
pcl, pcl_grad = [], []
for q in q_raw:
    q = q.clone().requires_grad_(True)
    val = ns["point_constraint_loss"](q, source, target)
    val.backward()
    pcl.append(val.detach())
    pcl_grad.append(q.grad.clone())
np.savez_compressed(
    os.path.join(HERE, "views.npz"), q1=q1.numpy(), q2=q2.numpy(), points=pts.numpy(),
    multiply=qu.quaternion_multiply(q1, q2).numpy(), apply=qu.quaternion_apply(q1, pts).numpy(),
    invert=qu.quaternion_invert(q1).numpy(), camera_positions=cam_p.numpy(), camera_orientations=cam_q.numpy(),
    q_raw=q_raw.numpy(), apply_raw=qu.quaternion_apply(q_raw, pts[:6]).numpy(), source=source.numpy(),
    target=target.numpy(), constraint_loss=torch.stack(pcl).numpy(), constraint_grad=torch.stack(pcl_grad).numpy(),
    position=pos.numpy(), orientation=ori.numpy(), position_c=torch.stack(pos_c).numpy(),
    orientation_c=torch.stack(ori_c).numpy())
print("views.npz written")
