"""CPU oracle for the SDF depth renderer -- TEST INFRASTRUCTURE ONLY.

Nothing under ``sdfest_b200/`` imports this package.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``; the measurement scripts under ``scripts/`` (not shipped, not imported by
the package) load ``oracle/_ref`` -- the reference's own extension -- only to time it beside
the product's kernels.  See ``oracle/sdf_oracle.c`` for the parity status.
"""
from .oracle import (  # noqa: F401
    build,
    composite_min_depth,
    l1_depth_loss,
    render,
    render_backward,
    render_composite,
    render_composite_backward,
)
