"""B200-native conditional RealNVP (forward / inverse / log-det) behind the reference's Python surface.

    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    from arl_conditional_normalizing_flows_b200.TOYcINN_make_model import cINN_affine

Adding this package directory to PYTHONPATH also exposes the reference's own module names
(`conv_cINN_make_model`, `conv_cINN_base_functions`, `TOYcINN_make_model`); see INTEGRATION.md.
Importing the package loads libcnf.so (hand-written sm_100a kernels); there is no fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if libcnf.so is missing)

__all__ = ["conv_cINN_make_model", "conv_cINN_base_functions", "TOYcINN_make_model", "initializers"]
__version__ = "0.1.0"
