"""Drop-in for ``src/barrier_method.py``: gradient descent on  f + beta*sum(reg) + beta*t*sum(phi),
phi = -log(-g(x)), with t shrinking by ``t_step`` per outer iteration and a projection onto the feasible
set after every step.  Same signature, same return value ``(x, opt_stats)``.

The driver is array-library agnostic: it only uses ``+ - *``, ``.copy()``/``.clone()`` and reductions that
both numpy arrays and torch (CUDA) tensors provide, so the closures may keep everything on the device (see
``whitomo_b200.solvers`` for the fused device loops that implement the reference's two concrete scenarios
with K FP + K BP per inner iteration instead of 3K FP + K BP).

Py3 fixes behind the same behaviour: ``.items()`` for ``.iteritems()`` (B3), barriers
without a projection are skipped via ``project`` (B4).  Kept on purpose: the gradient returned by
``goal_function[1]`` is modified in place (B5); the stats callback fires on every inner iteration because
``glob_iter`` only advances per outer iteration (B6); ``n_iter`` grows only under ``do_alpha_decay`` (B7).
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
    """sum over all elements as a python float / 0-d value (np.sum(np.ravel(.)) in the reference)."""
    if _is_tensor(x):
        return x.sum().item()
    return np.sum(np.ravel(x))


class BarrierFunction(object):
    """phi(x) = -log(-f(x)) for the constraint f(x) <= 0  (src/barrier_method.py:14-77)."""

    def __init__(self, func, grad, proj=None):
        self.f = func
        self.g = grad
        self.num_feval = 0
        self.num_geval = 0
        self.proj = proj

    def eval(self, x, do_grad=True):
        fx = self.f(x)
        if _is_tensor(fx):
            phi = -torch.log(-fx)
            phi = torch.where(fx >= 0, torch.full_like(phi, float('inf')), phi)
        else:
            with np.errstate(invalid='ignore', divide='ignore'):
                phi = -np.log(-fx)
            phi[fx >= 0] = np.inf          # infeasible points: inf instead of nan
        self.last_f = fx
        self.last_phi = phi
        self.num_feval += 1
        grad_phi = None
        if do_grad:
            gx = self.g(x)
            grad_phi = -gx / fx
            self.last_g = gx
            self.last_grad = grad_phi
            self.num_geval += 1
        return (phi, grad_phi)

    def func(self, x):
        return self.eval(x, do_grad=False)[0]

    def grad(self, x):
        return self.eval(x)[1]

    def is_feasible(self, x):
        return bool((self.f(x) <= 0).all())

    def project(self, x):
        if self.proj is None:
            return x
        return self.proj(x)


def barrier_method(x0, goal_function, reg_dict={}, ineq_dict={}, n_iter=200, t0=1.0, n_biter=10, alpha=0.1,
                   beta_reg=1.0, t_step=0.1,
                   add_stat_cb=None, max_iter=1000,
                   do_alpha_decay=True,
                   min_delta_loss=1e-3,
                   n_steps_without_progress=None):
    """Minimise ``goal_function`` = (f, grad f) with regularisers ``reg_dict`` = {name: (r, grad r)} subject to
    ``ineq_dict`` = {name: (g, grad g[, proj]) | BarrierFunction} (constraints g(x) <= 0).
    Parameters and semantics as in src/barrier_method.py:79-116."""
    bf_objs = collections.OrderedDict()
    bf = collections.OrderedDict()
    for k, v in ineq_dict.items():
        if type(v) is tuple:
            v = BarrierFunction(*v)
        elif not isinstance(v, BarrierFunction):
            continue
        bf_objs[k] = v
        bf[k] = (v.func, v.grad)

    # the initial guess must be feasible; one projection attempt, else (x0, None) like the reference
    if not all(v.is_feasible(x0) for v in bf_objs.values()):
        print('WARNING! the initial guess is not feasible.')
        print('trying to project in onto fieasible set..')
        for v in bf_objs.values():
            x0 = v.project(x0)
        infeas_list = [k for k, v in bf_objs.items() if not v.is_feasible(x0)]
        if infeas_list:
            print('ERROR! feasibility projection unsuccesfull')
            print('non-satisfied constraints: ', ' '.join(infeas_list))
            return x0, None

    t = t0
    x = x0
    opt_stats = []
    glob_iter = 0
    for i_biter in range(n_biter):
        prev_loss = None
        no_progress_count = 0
        print('starting ', i_biter, 'barrier iteration.')
        for i_iter in range(n_iter):
            goal_grad = goal_function[1](x)
            reg_grads = collections.OrderedDict((k, v[1](x)) for k, v in reg_dict.items())
            bf_grads = collections.OrderedDict((k, v[1](x)) for k, v in bf.items())

            grad = goal_grad                      # same object: accumulated in place (B5)
            for g in reg_grads.values():
                grad += beta_reg * g
            for g in bf_grads.values():
                grad += beta_reg * g * t
            step = -alpha * grad

            helpers.printProgressBar(i_iter, n_iter - 1,
                                     prefix='barrier iter %d/%d' % (i_biter + 1, n_biter),
                                     suffix='done (%03d/%03d)' % (i_iter + 1, n_iter), length=50)

            if add_stat_cb is not None and (glob_iter % 10 == 0 or glob_iter < 30):
                x_stat = _copy(x)
                goal_stat = StepStat(x_stat, goal_function[0](x), _copy(goal_grad))
                reg_stats = {k: StepStat(x_stat, v[0](x), _copy(reg_grads[k])) for k, v in reg_dict.items()}
                bf_stats = {k: StepStat(x_stat, _copy(v.last_phi), _copy(v.last_grad)) for k, v in bf_objs.items()}
                add_stat_cb((goal_stat, reg_stats, bf_stats))

            x = x + step
            for v in bf_objs.values():            # projection after every step
                x = v.project(x)

            if n_steps_without_progress is not None:
                loss = goal_function[0](x)
                for v in reg_dict.values():
                    loss += _total(beta_reg * v[0](x))
                for v in bf.values():
                    loss += _total(beta_reg * t * v[0](x))
                if prev_loss is not None:
                    if prev_loss - loss >= min_delta_loss:
                        no_progress_count = 0
                    else:
                        no_progress_count += 1
                prev_loss = loss
                if no_progress_count >= n_steps_without_progress:
                    print('\nstopping barrier step', i_biter, 'on iteration', i_iter,
                          'because there was no valuable progress')
                    break

        t = t * t_step
        if do_alpha_decay:
            alpha = alpha * math.sqrt(t_step)
            n_iter = int(n_iter / t_step)
            if n_iter > max_iter:
                n_iter = max_iter
        print('')
        glob_iter += 1

    return x, opt_stats
