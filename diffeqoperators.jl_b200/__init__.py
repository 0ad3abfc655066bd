"""diffeqoperators.jl_b200 -- B200-native `mul!` / `*` for DiffEqOperators.jl-style operators.

Host side: a mirror of the reference's operator/BC constructors (same names, argument meaning and
error behaviour; Julia's `Ctor{N}(...)` is `Ctor[N](...)`).  Compute: libdeo_b200.so, hand-written
sm_100a CUDA behind the C ABI in include/deo_b200.h.  There is no CPU compute path in this package.

The directory name contains a dot, so import it through the repo-root shim:  `import deo_b200`.
"""
from ._lib import DeoError, LIB_PATH, launch_count, load as load_library
from .operators import (CenteredDifference, UpwindDifference, DerivativeOperator, GhostDerivativeOperator,
                        DiffEqOperatorCombination, Laplacian, calculate_weights)
from .bc import (RobinBC, GeneralBC, NeumannBC, DirichletBC, Dirichlet0BC, Neumann0BC, PeriodicBC,
                 MultiDimBC, MultiDimDirectionalBC, ComposedMultiDimBC, AffineBC, BoundaryPadded, compose)
from .device import DeviceArray, zeros, sync
from .apply import mul_, mul_alloc, step_, Plan, build_plans
from .vector_calculus import (Gradient, Divergence, Curl, GradientOperator, DivergenceOperator, CurlOperator,
                              nonlinear_diffusion, nonlinear_diffusion_, DiffEqOperatorComposition, compose_operators,
                              concretize, ldiv)

__all__ = [n for n in dir() if not n.startswith("_")]
