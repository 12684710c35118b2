"""sdfest_b200 -- the differentiable SDF depth renderer of roym899/sdfest, rebuilt for B200.

Only the hot path lives here: ``sdfest_b200.differentiable_renderer`` (drop-in for
``sdfest.differentiable_renderer``) on top of ``libsdfrender.so`` (hand-written sm_100a kernels,
C ABI in ``include/sdfrender.h``), plus the batched render-and-compare loop that drives it
(``sdfest_b200.estimation``).  Build the library with ``python -m sdfest_b200.build``.
"""
__version__ = "0.1.0"
