"""Host mathematics of the multi-view loop (sdfest_b200/estimation/views.py) on CPU: against golden
vectors of the reference's own quaternion_utils / view-loop lines, and the analytic pull-back of the
pose gradients against autograd."""
import os

import numpy as np
import torch

from sdfest_b200.estimation import views
from util import GOLDEN_DIR


def _golden():
    z = np.load(os.path.join(GOLDEN_DIR, "views.npz"))
    return {k: torch.tensor(z[k]) for k in z.files}


def test_quaternion_helpers_match_reference_golden():
    z = _golden()
    torch.testing.assert_close(views.quaternion_multiply(z["q1"], z["q2"]), z["multiply"], rtol=0, atol=1e-14)
    torch.testing.assert_close(views.quaternion_apply(z["q1"], z["points"]), z["apply"], rtol=0, atol=1e-14)
    assert torch.equal(views.quaternion_invert(z["q1"]), z["invert"])
    # broadcasting as the reference's ("normal broadcasting rules apply")
    out = views.quaternion_multiply(z["q1"][:, None], z["q2"][None])
    assert out.shape == (7, 7, 4)
    torch.testing.assert_close(out[2, 5], views.quaternion_multiply(z["q1"][2], z["q2"][5]))


def test_camera_frames_match_the_reference_view_loop():
    z = _golden()
    pos_c, ori_c = views.to_camera_frames(z["position"], z["orientation"], z["camera_positions"],
                                          z["camera_orientations"])
    torch.testing.assert_close(pos_c, z["position_c"], rtol=0, atol=1e-14)
    torch.testing.assert_close(ori_c, z["orientation_c"], rtol=0, atol=1e-14)
    # an identity camera at the origin leaves the pose alone
    eye_p, eye_q = torch.zeros(1, 3, dtype=torch.float64), torch.tensor([[0.0, 0, 0, 1]], dtype=torch.float64)
    p1, q1 = views.to_camera_frames(z["position"], z["orientation"], eye_p, eye_q)
    assert torch.equal(p1[0], z["position"]) and torch.equal(q1[0], z["orientation"])


def test_pull_back_is_the_adjoint_of_the_view_maps():
    z = _golden()
    pos = z["position"].clone().requires_grad_(True)
    ori = z["orientation"].clone().requires_grad_(True)
    pos_c, ori_c = views.to_camera_frames(pos, ori, z["camera_positions"], z["camera_orientations"])
    g = torch.Generator().manual_seed(1)
    g_pc = torch.randn(pos_c.shape, generator=g, dtype=torch.float64)
    g_qc = torch.randn(ori_c.shape, generator=g, dtype=torch.float64)
    ((pos_c * g_pc).sum() + (ori_c * g_qc).sum()).backward()
    g_p, g_q = views.pull_back(g_pc, g_qc, z["camera_orientations"])
    torch.testing.assert_close(g_p, pos.grad, rtol=0, atol=1e-13)
    torch.testing.assert_close(g_q, ori.grad, rtol=0, atol=1e-13)
