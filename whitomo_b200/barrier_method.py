"""Drop-in for ``src/barrier_method.py``: gradient descent on  f + beta*sum(reg) + beta*t*sum(phi),
phi = -log(-g(x)), with t shrinking by ``t_step`` per outer iteration and a projection onto the feasible
set after every step.  Same public names (``StepStat``, ``BarrierFunction``, ``barrier_method``),
same signature, same return value ``(x, opt_stats)``.

The driver is array-library agnostic: the closures may work on numpy arrays or on torch (CUDA) tensors, so a
caller can keep the whole loop on the device (``whitomo_b200.solvers`` holds the fused device loops for the
reference's two concrete scenarios, which spend K FP + K BP per inner iteration instead of 3K FP + K BP).

Behaviour kept on purpose: the array returned by the goal gradient is accumulated into in
place, so the goal ``StepStat.grad`` is the total gradient (B5); the statistics callback fires on every inner
iteration because the counter that gates it only advances per outer iteration (B6); the inner iteration count
grows only together with the step decay (B7).  Fixed: dict iteration is py3 (B3), a barrier without a
projection is skipped instead of raising (B4).
"""
from __future__ import print_function, division

import collections
import math

import numpy as np

from . import helpers

try:  # torch is only needed when the caller passes tensors
    import torch
except ImportError:  # pragma: no cover
    torch = None

StepStat = collections.namedtuple('StepStat', ['x', 'func', 'grad'])


def _is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _copy(x):
    return x.clone() if _is_tensor(x) else x.copy()


def _total(x):
    """sum over all elements, as the reference's np.sum(np.ravel(.))"""
    return x.sum().item() if _is_tensor(x) else np.sum(np.ravel(x))


def _neg_log_barrier(fx):
    """-log(-f) with +inf (not nan) wherever the constraint f <= 0 is violated or active"""
    if _is_tensor(fx):
        phi = -torch.log(-fx)
        return torch.where(fx >= 0, torch.full_like(phi, float('inf')), phi)
    with np.errstate(invalid='ignore', divide='ignore'):
        phi = -np.log(-fx)
    phi[fx >= 0] = np.inf
    return phi


class BarrierFunction(object):
    """Barrier penalty phi(x) = -log(-f(x)) of one constraint f(x) <= 0 (src/barrier_method.py:14-77).

    ``func`` / ``grad`` evaluate f and its gradient, ``proj`` (optional) maps a point onto the feasible set.
    ``eval`` caches ``last_f``, ``last_phi``, ``last_g``, ``last_grad`` for the statistics callback and counts
    evaluations in ``num_feval`` / ``num_geval`` like the reference."""

    def __init__(self, func, grad, proj=None):
        self.f, self.g, self.proj = func, grad, proj
        self.num_feval = self.num_geval = 0

    def eval(self, x, do_grad=True):
        self.last_f = self.f(x)
        self.last_phi = _neg_log_barrier(self.last_f)
        self.num_feval += 1
        if not do_grad:
            return (self.last_phi, None)
        self.last_g = self.g(x)
        self.last_grad = -self.last_g / self.last_f
        self.num_geval += 1
        return (self.last_phi, self.last_grad)

    def func(self, x):
        return self.eval(x, do_grad=False)[0]

    def grad(self, x):
        return self.eval(x)[1]

    def is_feasible(self, x):
        return bool((self.f(x) <= 0).all())

    def project(self, x):
        return x if self.proj is None else self.proj(x)


def _as_barriers(ineq_dict):
    """{name: (g, grad g[, proj]) | BarrierFunction} -> ordered {name: BarrierFunction}"""
    out = collections.OrderedDict()
    for name, spec in ineq_dict.items():
        if type(spec) is tuple:
            out[name] = BarrierFunction(*spec)
        elif isinstance(spec, BarrierFunction):
            out[name] = spec
    return out


def _make_feasible(x0, barriers):
    """src/barrier_method.py:128-152: one projection attempt; returns (x0, ok)"""
    if all(b.is_feasible(x0) for b in barriers.values()):
        return x0, True
    print('WARNING! the initial guess is not feasible.')
    print('trying to project in onto fieasible set..')
    for b in barriers.values():
        x0 = b.project(x0)
    failed = [name for name, b in barriers.items() if not b.is_feasible(x0)]
    if failed:
        print('ERROR! feasibility projection unsuccesfull')
        print('non-satisfied constraints: ', ' '.join(failed))
    return x0, not failed


def _penalised_loss(x, goal_function, reg_dict, barriers, beta_reg, t):
    """f(x) + sum beta*reg(x) + sum beta*t*phi(x), every term summed over all elements (:208-213)"""
    loss = goal_function[0](x)
    for r in reg_dict.values():
        loss += _total(beta_reg * r[0](x))
    for b in barriers.values():
        loss += _total(beta_reg * t * b.func(x))
    return loss


def barrier_method(x0, goal_function, reg_dict={}, ineq_dict={}, n_iter=200, t0=1.0, n_biter=10, alpha=0.1,
                   beta_reg=1.0, t_step=0.1,
                   add_stat_cb=None, max_iter=1000,
                   do_alpha_decay=True,
                   min_delta_loss=1e-3,
                   n_steps_without_progress=None):
    """Minimise ``goal_function`` = (f, grad f) plus regularisers ``reg_dict`` = {name: (r, grad r)} subject to the
    inequality constraints of ``ineq_dict`` (g(x) <= 0).  Parameters as in src/barrier_method.py:84-116:
    ``n_biter`` barrier levels starting at ``t0`` and shrinking by ``t_step``; ``n_iter`` gradient steps of size
    ``alpha`` per level; with ``do_alpha_decay`` the step shrinks by sqrt(t_step) and the number of steps grows by
    1/t_step (capped at ``max_iter``) per level; a level is left early after ``n_steps_without_progress`` steps
    whose loss decrease stays below ``min_delta_loss``.  Returns ``(x, [])`` or ``(x0, None)`` when the start
    cannot be made feasible."""
    barriers = _as_barriers(ineq_dict)
    x, ok = _make_feasible(x0, barriers)
    if not ok:
        return x, None

    opt_stats = []
    t = t0
    stats_gate = 0            # the reference's glob_iter: advanced once per barrier level
    for level in range(n_biter):
        print('starting ', level, 'barrier iteration.')
        last_loss, stalled = None, 0
        for step_no in range(n_iter):
            total = goal_function[1](x)            # accumulated into in place below (B5)
            reg_parts = collections.OrderedDict((name, r[1](x)) for name, r in reg_dict.items())
            bar_parts = collections.OrderedDict((name, b.grad(x)) for name, b in barriers.items())
            goal_part = total
            for part in reg_parts.values():
                total += beta_reg * part
            for part in bar_parts.values():
                total += beta_reg * part * t

            helpers.printProgressBar(step_no, n_iter - 1, prefix='barrier iter %d/%d' % (level + 1, n_biter),
                                     suffix='done (%03d/%03d)' % (step_no + 1, n_iter), length=50)

            if add_stat_cb is not None and (stats_gate % 10 == 0 or stats_gate < 30):
                snap = _copy(x)
                add_stat_cb((StepStat(snap, goal_function[0](x), _copy(goal_part)),
                             {name: StepStat(snap, r[0](x), _copy(reg_parts[name])) for name, r in reg_dict.items()},
                             {name: StepStat(snap, _copy(b.last_phi), _copy(b.last_grad)) for name, b in barriers.items()}))

            x = x + (-alpha * total)
            for b in barriers.values():            # projection after every step (:203-206)
                x = b.project(x)

            if n_steps_without_progress is None:
                continue
            loss = _penalised_loss(x, goal_function, reg_dict, barriers, beta_reg, t)
            if last_loss is not None:
                stalled = 0 if last_loss - loss >= min_delta_loss else stalled + 1
            last_loss = loss
            if stalled >= n_steps_without_progress:
                print('\nstopping barrier step', level, 'on iteration', step_no,
                      'because there was no valuable progress')
                break

        t *= t_step
        if do_alpha_decay:
            alpha *= math.sqrt(t_step)
            n_iter = min(int(n_iter / t_step), max_iter)
        print('')
        stats_gate += 1

    return x, opt_stats
