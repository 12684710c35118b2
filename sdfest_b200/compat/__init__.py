"""Binary-compatible stand-ins for pieces of the reference that this library replaces."""
from . import sdf_renderer_cpp  # noqa: F401
