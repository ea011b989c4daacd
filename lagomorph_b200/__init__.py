"""lagomorph_b200 -- Blackwell-native LDDMM geodesic-shooting hot path.

Drop-in for the hot-path operators of jacobhinkle/lagomorph (same names and
argument meaning as `import lagomorph as lm`), backed by hand-written sm_100a
CUDA kernels behind the C ABI declared in include/lagomorph_b200.h.
"""
from . import _lib
from .deform import (identity, interp, interp_adjoint, interp_hessian_diagonal_image, compose,
                     compose_disp_vel, compose_vel_disp, InterpFunction)
from .diff import (jacobian_times_vectorfield, jacobian_times_vectorfield_adjoint,
                   JacobianTimesVectorFieldFunction, JacobianTimesVectorFieldAdjointFunction)
from .adjrep import ad, Ad, ad_star, Ad_star, coad, ad_dagger, Ad_dagger, sym, sym_dagger
from .metric import FluidMetric, FluidMetricOperator, Metric, fluid_operator
from .lddmm import expmap, expmap_advect, expmap_host, EPDiff_step, EPDiffStep, EPDiff_steps
from .affine import (regrid, RegridFunction, RegridModule, affine_interp, AffineInterp,
                     AffineInterpFunction, affine_inverse, det_2x2)
from .atlas import LDDMMAtlasBuilder, lddmm_atlas
from .affine_atlas import affine_atlas, StandardizedDataset, save_affine_atlas, load_affine_atlas

__version__ = "0.1.0"


def set_debug_mode(mode):
    """Synchronise and raise CUDA errors after every library call (reference:
    lagomorph_ext.set_debug_mode, which only printed them)."""
    _lib.lib.lgm_set_debug_mode(1 if mode else 0)


def launch_count():
    """Number of kernels this library has launched in this process."""
    return int(_lib.lib.lgm_launch_count())
