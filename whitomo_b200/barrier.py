"""Drop-in for ``teeth_recon/barrier.py``: ``HoughBarrier``, the log-barrier on the sinogram ("Hough") domain --
the constraint FP(x)[m] >= b for the photon-starved rays ``m`` of a metal shadow, i.e. f(x) = b - FP(x)[m] <= 0.

``FP`` / ``BP`` are the caller's closures (flat in, flat out) exactly as in the reference; they may be numpy
closures over ``astra_proxy.gpu_fp`` or torch closures over ``Projector.fp`` -- the class works on either array
type.  The fused device loop for the mono MAR scenario lives in ``whitomo_b200.solvers``.
"""
import numpy as np

from . import barrier_method
from .barrier_method import _is_tensor, _copy, _neg_log_barrier

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


class HoughBarrier(barrier_method.BarrierFunction):
    #: Landweber-like feasibility projection (teeth_recon/barrier.py:30): at most this many steps of this size
    PROJECT_STEPS = 100
    PROJECT_STEP_SIZE = 1e-2

    def __init__(self, FP, BP, m, b, r_like):
        self.FP, self.BP, self.m, self.b, self.r = FP, BP, m, b, r_like
        super(HoughBarrier, self).__init__(self._constraint, self._constraint_grad, self._project_feasible)

    # ---- the (func, grad, proj) triple BarrierFunction expects --------------------------------------------
    def _sino_zeros(self):
        z = torch.zeros_like(self.r) if _is_tensor(self.r) else np.zeros_like(self.r)
        return z.flatten()

    def _constraint(self, x):
        return self.b - self.FP(x)[self.m]

    def _constraint_grad(self, x):
        ones_on_mask = self._sino_zeros()          # the reference reads an undefined name here (B10)
        ones_on_mask[self.m] = -1
        return self.BP(ones_on_mask)

    def _project_feasible(self, x, n_iter=None, alpha=None):
        """Push x towards {FP(x)[m] >= b}: repeat x -= alpha * BP(residual of the violated rays), where the
        residual is FP(x) - (b + 1e-3) on the violated rays and 0 elsewhere (the reference's literal update,
        teeth_recon/barrier.py:36-43, B11)."""
        n_iter = self.PROJECT_STEPS if n_iter is None else n_iter
        alpha = self.PROJECT_STEP_SIZE if alpha is None else alpha
        proj = self.FP(x)
        out = _copy(x)
        for _ in range(n_iter):
            violated = self.m & (proj < self.b)
            if not bool(violated.any()):
                break
            proj[~violated] = 0
            proj[violated] -= self.b + 1e-3
            out = out - alpha * self.BP(proj)
            proj = self.FP(out)
        return out

    # ---- evaluation with ONE forward and ONE back projection (teeth_recon/barrier.py:50-69) ----------------
    def eval(self, x, do_grad=True):
        self.last_f = self.b - self.FP(x)[self.m]
        self.last_phi = _neg_log_barrier(self.last_f)
        self.num_feval += 1
        if not do_grad:
            return (self.last_phi, None)
        weights = self._sino_zeros()
        weights[self.m] = -1 / self.last_f
        self.last_g = weights
        self.last_grad = -self.BP(weights)
        self.num_geval += 1
        return (self.last_phi, self.last_grad)
