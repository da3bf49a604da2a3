"""faer_amg_b200 -- B200-native (sm_100a) implementation of the faer-amg hot path.

Python here is host plumbing only: every class is a thin handle over the C ABI of ``libfamg.so``
(``include/famg.h``), mirroring the reference crate's module and type names
(``core``, ``par_spmm``, ``hierarchy``, ``interpolation``, ``preconditioners::{multigrid, smoothers,
coarse_solvers, block_smoothers}``).  There is no CPU fallback.
"""
from . import _ffi  # noqa: F401
from ._ffi import FamgError  # noqa: F401
from .core import Context, DeviceMat, ParSpmmOp, SparseMatOp, SparseRowMat, PAR_BLOCK_SIZE  # noqa: F401
from .hierarchy import Hierarchy, HierarchyConfig  # noqa: F401
from .interpolation import (AggregationConfig, GalerkinCoarse, InterpolationConfig, block_jacobi, galerkin_product,  # noqa: F401
                            smooth_interpolation, smooth_p, smoothed_aggregation, tentative_prolongator)
from .partitioners import (GeometricPartitioner, Partition, PartitionerConfig, StrengthGraph,  # noqa: F401
                           geometric_partition)
from . import partitioners  # noqa: F401
from .preconditioners.block_smoothers import BlockSmoother, BlockSmootherConfig  # noqa: F401
from .preconditioners.coarse_solvers import CoarseSolverKind, SparseCholeskySolve  # noqa: F401
from .preconditioners.multigrid import Multigrid, MultigridConfig  # noqa: F401
from .preconditioners.composite import Composite  # noqa: F401
from .preconditioners.smoothers import (Diag, SmootherKind, StationaryIteration, new_jacobi, new_l1, new_l2,  # noqa: F401
                                        smooth)
from .solvers import (CgError, CgInfo, CgParams, conjugate_gradient, conjugate_gradient_dev, stationary_solver,  # noqa: F401
                      test_solver)
from .adaptivity import (AdaptiveConfig, ErrorPropogator, create_weights, find_near_null, smooth_vector,  # noqa: F401
                         smooth_vector_dev)
from . import adaptivity  # noqa: F401
from . import gallery  # noqa: F401
from .utils import approx_convergence_factor, mats_are_equal, symmetry_test  # noqa: F401

__version__ = "0.1.0"
