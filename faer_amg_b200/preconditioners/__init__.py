"""``src/preconditioners/`` of the reference: multigrid, smoothers, coarse solvers, block smoothers."""
from . import block_smoothers, coarse_solvers, multigrid, smoothers  # noqa: F401
