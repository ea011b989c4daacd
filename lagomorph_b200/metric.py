"""Fluid and other LDDMM metrics (mirror of lagomorph/metric.py).

sharp/flat are one library call each: own batched R2C/C2R FFT passes with the
Fourier multiplier fused into the middle pass (csrc/fluid.cu), instead of the
reference's torch.rfft -> lagomorph_ext.fluid_operator -> torch.irfft triple.
"""
import ctypes

import numpy as np
import torch

from . import _lib as L


def fluid_apply(params, mv, inverse, out=None):
    """F^-1[ L(k)^-1 F[mv] ] (inverse=True, sharp) or F^-1[ L(k) F[mv] ] (flat)."""
    dev = L.require_cuda(mv)
    d = L.spatial_dim(mv)
    if mv.shape[1] != d:
        raise RuntimeError("Vector field has incorrect shape for dimension")
    alpha, beta, gamma = [float(p) for p in params]
    mv = L.aligned(mv)
    if out is None:
        out = torch.empty_like(mv)
    N = mv.shape[0]
    sh = L.shape_arr(mv.shape[2:])
    code = L.dtype_code(mv)
    nbytes = L.lib.lgm_fluid_workspace_bytes(code, N, d, sh)
    if nbytes < 0:
        raise RuntimeError("lgm_fluid_workspace_bytes rejected the arguments")
    ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_fluid_apply(code, L.ptr(out), L.ptr(mv), N, d, sh, int(bool(inverse)), alpha, beta,
                                      gamma, L.ptr(ws), int(nbytes), L.stream_ptr(dev)))
    return out


def fluid_operator(Fmv, inverse, cosluts, sinluts, alpha, beta, gamma):
    """In-place Fourier multiplier on an interleaved half spectrum (N,d,X,Y[,Zc],2) from a
    unitary rfftn: the reference's lagomorph_ext.fluid_operator (extension.cpp:158-173)."""
    dev = L.require_cuda(Fmv, *cosluts, *sinluts)
    d = Fmv.dim() - 3
    if d not in (2, 3):
        raise RuntimeError("Only two- and three-dimensional fluid metric is supported")
    if Fmv.shape[1] != d:
        raise RuntimeError("Vector field has incorrect shape for dimension")
    if not Fmv.is_contiguous():
        raise RuntimeError("Fmv must be contiguous")
    cl = (ctypes.c_void_p * 3)(*([c.data_ptr() for c in cosluts] + [None] * (3 - d)))
    sl = (ctypes.c_void_p * 3)(*([s.data_ptr() for s in sinluts] + [None] * (3 - d)))
    with torch.cuda.device(dev):
        L.check(L.lib.lgm_fluid_operator(L.dtype_code(Fmv), L.ptr(Fmv), int(bool(inverse)), cl, sl, float(alpha),
                                         float(beta), float(gamma), Fmv.shape[0], d,
                                         L.shape_arr(Fmv.shape[2:2 + d]), L.stream_ptr(dev)))


class FluidMetricOperator(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, luts, inverse, mv):
        ctx.params = params
        ctx.inverse = inverse
        return fluid_apply(params, mv, inverse)

    @staticmethod
    def backward(ctx, outgrad):
        # the operator is linear and symmetric: backward == forward (reference: metric.py:21-34)
        return None, None, None, fluid_apply(ctx.params, outgrad, ctx.inverse)


class FluidMetric(object):
    def __init__(self, params=[0.1, 0.0, 0.001]):
        """Green's function metric for L'L = -alpha lap - beta grad div + gamma
        (reference: metric.py:37-51)."""
        self.shape = None
        self.complexshape = None
        assert len(params) == 3
        self.params = params
        self.luts = None

    def initialize_luts(self, shape, dtype, device="cuda"):
        """cos/sin lookup tables exactly as the reference builds them (metric.py:53-75), except
        that they are rebuilt when the shape, dtype or device changes (the reference never does).
        The fused sharp/flat path builds its own copies inside the library; these are for
        fluid_operator() users."""
        key = (tuple(shape), dtype, torch.device(device))
        if getattr(self, "_lut_key", None) == key and self.luts is not None:
            return
        self._lut_key = key
        self.shape = shape
        self.complexshape = list(shape)
        self.complexshape[-1] = self.complexshape[-1] // 2 + 1
        self.complexshape = tuple(self.complexshape)
        self.luts = {"cos": [], "sin": []}
        for (Nf, N) in zip(self.complexshape[2:], self.shape[2:]):
            self.luts["cos"].append(
                torch.Tensor(2.0 * (1.0 - np.cos(2 * np.pi * np.arange(Nf) / N))).type(dtype).to(device))
            self.luts["sin"].append(
                torch.Tensor(np.sin(2.0 * np.pi * np.arange(Nf) / N)).type(dtype).to(device))

    def operator(self, mv, inverse):
        # like the reference (metric.py:78), so that .luts / .shape / .complexshape are there for code
        # that reads them; cached per (shape, dtype, device) and skipped under CUDA-graph capture
        if mv.is_cuda and not torch.cuda.is_current_stream_capturing():
            self.initialize_luts(shape=mv.shape, dtype=mv.dtype, device=mv.device)
        return FluidMetricOperator.apply(self.params, self.luts, inverse, mv)

    def sharp(self, m):
        """Momentum -> velocity: apply the Green's function (reference: metric.py:81-88)."""
        return self.operator(m, inverse=True)

    def flat(self, m, out=None):
        """Velocity -> momentum: apply the differential operator (reference: metric.py:90-97)."""
        return self.operator(m, inverse=False)


class Metric:
    """Serialization and command line interface to a metric factory (reference: metric.py:100-135)."""

    @staticmethod
    def add_args(parser):
        parser.add_argument("--metric_type", default="fluid", type=str,
                            help="Type of parser. Currently only 'fluid' is supported.")
        parser.add_argument("--fluid_alpha", default=0.1, type=float,
                            help="Fluid parameter for vector Laplacian term")
        parser.add_argument("--fluid_beta", default=0.0, type=float,
                            help="Fluid parameter for gradient divergence term")
        parser.add_argument("--fluid_gamma", default=0.01, type=float, help="Fluid parameter for L2 term")

    @classmethod
    def from_args(cls, args):
        if args.metric_type.lower() == "fluid":
            return FluidMetric(params=[args.fluid_alpha, args.fluid_beta, args.fluid_gamma])
