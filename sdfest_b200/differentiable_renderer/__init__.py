"""Differentiable depth renderer for discretised SDFs, B200-native.

Mirrors ``sdfest.differentiable_renderer`` (whose package exports ``Camera`` and
``render_depth_gpu``, reference ``__init__.py:6-8``) and adds the batched entry points.
"""
from .sdf_renderer import (  # noqa: F401
    Camera,
    SDFRendererFunctionGPU,
    forward_stats,
    get_empty_space_policy,
    grid_bounds,
    get_sdf_grad_mode,
    get_sdf_layout_policy,
    render_and_compare,
    render_depth,
    render_depth_batched,
    render_depth_composite,
    render_depth_gpu,
    set_empty_space_policy,
    set_sdf_grad_mode,
    set_sdf_layout_policy,
)
