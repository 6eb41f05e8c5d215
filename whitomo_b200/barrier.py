"""Drop-in for ``teeth_recon/barrier.py``: ``HoughBarrier``, the log-barrier on the sinogram domain
(constraint FP(x)[m] >= b for the photon-starved rays of a metal shadow).

``FP`` / ``BP`` are the caller's closures (flat in, flat out) exactly as in the reference; they may be
numpy closures over ``astra_proxy.gpu_fp`` or torch closures over ``Projector.fp`` -- the class works on
either array type.  The fused device loop for the mono MAR scenario lives in ``whitomo_b200.solvers``.
"""
import numpy as np

from . import barrier_method
from .barrier_method import _is_tensor, _copy

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _zeros_like_flat(r):
    if _is_tensor(r):
        return torch.zeros_like(r).flatten()
    return np.zeros_like(r).flatten()


class HoughBarrier(barrier_method.BarrierFunction):
    def __init__(self, FP, BP, m, b, r_like):
        self.FP = FP
        self.BP = BP
        self.m = m
        self.b = b
        self.r = r_like

        def bc_func(x):
            return b - FP(x)[m]

        def bc_grad(x):
            grad = _zeros_like_flat(r_like)       # the reference reads an undefined name here (B10)
            grad[m] = -1
            return BP(grad)

        def bc_project(x, n_iter=100, alpha=1e-2):
            """teeth_recon/barrier.py:30-45 -- Landweber-like steps on the violated rays only (B11 kept)."""
            f = FP(x)
            x_new = _copy(x)
            for _ in range(n_iter):
                infeas_i = m & (f < b)
                if not bool(infeas_i.any()):
                    break
                f[~infeas_i] = 0
                f[infeas_i] -= b + 1e-3
                x_new = x_new - alpha * BP(f)
                f = FP(x_new)
            return x_new

        super(HoughBarrier, self).__init__(bc_func, bc_grad, bc_project)

    def eval(self, x, do_grad=True):
        fp = self.FP(x)
        fx = self.b - fp[self.m]
        if _is_tensor(fx):
            phi = -torch.log(-fx)
            phi = torch.where(fx >= 0, torch.full_like(phi, float('inf')), phi)
        else:
            with np.errstate(invalid='ignore', divide='ignore'):
                phi = -np.log(-fx)
            phi[fx >= 0] = np.inf
        self.last_f = fx
        self.last_phi = phi
        self.num_feval += 1
        grad_phi = None
        if do_grad:
            grad = _zeros_like_flat(self.r)
            grad[self.m] = -1 / fx
            grad_phi = -self.BP(grad)
            self.last_g = grad
            self.last_grad = grad_phi
            self.num_geval += 1
        return (phi, grad_phi)
