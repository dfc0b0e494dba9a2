"""cvxpnpl_b200 -- B200-native batched CvxPnPL (drop-in for cvxpnpl's pnp/pnl/pnpl).

Scalar drop-in API (reference signatures): pnp, pnl, pnpl, CvxPnPL.
Batched API (torch CUDA tensors with a leading batch dimension): pnp_batched,
pnl_batched, pnpl_batched, solve_batched.
"""
__version__ = "0.1.0"

from .api import CvxPnPL, null, pnl, pnp, pnpl, rc, rc_pnl, rc_pnp, rc_pnpl  # noqa: F401
from .batched import (BatchedPoses, Workspace, assemble_batched, extract_batched, pnl_batched,  # noqa: F401
                      pnp_batched, pnpl_batched, solve_batched, solve_sdp_batched, measure_fp64_peak, null_batched, last_kernel_times)
from .pipeline import HostPipeline, HostStager  # noqa: F401,E402
