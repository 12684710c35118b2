"""Host-side mirror of the reference interface: Camera conventions and the error behaviour of
render_depth_gpu (reference sdf_renderer.py:31-133, 360-424; sdf_renderer.cpp:9-13)."""
import math

import pytest
import torch

from sdfest_b200.differentiable_renderer import (Camera, get_sdf_grad_mode, render_depth,
                                                 render_depth_batched, render_depth_gpu,
                                                 set_sdf_grad_mode)
from sdfest_b200.differentiable_renderer import sdf_renderer as sr


def test_camera_pixel_center_conversion():
    cam = Camera(640, 480, 320, 320, 320, 240, pixel_center=0.5)  # default.yaml:1-8
    assert cam.get_pinhole_camera_parameters(0.5) == (320, 320, 320, 240, 0.0)
    assert cam.get_pinhole_camera_parameters(0.0) == (320, 320, 319.5, 239.5, 0.0)
    cam0 = Camera(640, 480, 525, 525, 319.5, 239.5, pixel_center=0.0)  # redwood.yaml:4-12
    fx, fy, cx, cy, s = cam0.get_pinhole_camera_parameters(0.5)
    assert (fx, fy, cx, cy, s) == (525, 525, 320.0, 240.0, 0.0)
    assert sr._camera_params(cam0) == (640, 480, 320.0, 240.0, 525.0, 525.0)


def test_fov_camera_matches_reference_formula():
    # sdf_renderer.py:418-420: f = W / tan(fov/2) / 2, principal point at the image centre
    W, H, fov = 100, 80, 90.0
    f = W / math.tan(fov * math.pi / 180.0 / 2.0) / 2
    assert f == pytest.approx(50.0)


def _cpu_args():
    return (torch.zeros(4, 4, 4), torch.zeros(3), torch.tensor([0, 0, 0, 1.0]), torch.ones(1))


def test_either_camera_or_fov():
    cam = Camera(8, 8, 4, 4, 4, 4, pixel_center=0.5)
    with pytest.raises(ValueError, match="Either width"):
        render_depth_gpu(*_cpu_args(), 8, 8, 60.0, 0.01, cam)
    with pytest.raises(ValueError, match="Either width"):
        render_depth_gpu(*_cpu_args())


def test_cpu_tensors_are_rejected_like_the_reference():
    with pytest.raises(RuntimeError, match="sdf must be a CUDA tensor"):
        render_depth_gpu(*_cpu_args(), 8, 8, 60.0, 0.01)
    cam = Camera(8, 8, 4, 4, 4, 4, pixel_center=0.5)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        render_depth_batched(torch.zeros(4, 4, 4), torch.zeros(2, 3), torch.zeros(2, 4),
                             torch.ones(2), 0.01, cam)


def test_no_cpu_renderer_in_the_product():
    with pytest.raises(NotImplementedError, match="no CPU renderer"):
        render_depth(*_cpu_args(), 8, 8, 60.0, 0.01)


def test_sdf_grad_mode_switch():
    assert get_sdf_grad_mode() == "reference"
    set_sdf_grad_mode("exact")
    try:
        assert get_sdf_grad_mode() == "exact"
        assert sr._grad_flags((True, False, True, False), None) == 0x01 | 0x04 | 0x10
        assert sr._grad_flags((True, True, True, True), "reference") == 0x0F
    finally:
        set_sdf_grad_mode("reference")
    with pytest.raises(ValueError):
        set_sdf_grad_mode("fast")
