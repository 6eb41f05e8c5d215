"""Oracle restatement of the log-barrier driver -- TEST INFRASTRUCTURE.

  BarrierFunction      src/barrier_method.py:14-77
  barrier_method       src/barrier_method.py:79-240   (SURVEY A.7 pseudo-code)
  HoughBarrier         teeth_recon/barrier.py:10-69
  ray_transform...     teeth_recon/experiments.py:290-293
  mono MAR closures    teeth_recon/experiments.py:651-749 (barrier_least_squares, first run)

numpy, py3 (``.items()`` instead of ``.iteritems()``).  Pinned by
tests/golden/barrier_*.npz, produced by executing the reference's own modules (py3-patched at load
time) on an oracle-backed ``astra_proxy`` -- see tests/golden/make_golden.py.
"""
import numpy as np


class BarrierFunction(object):
    """phi(x) = -log(-f(x)); grad phi = -g/f (src/barrier_method.py:14-77)."""

    def __init__(self, func, grad, proj=None):
        self.f, self.g, self.proj = func, grad, proj
        self.num_feval = self.num_geval = 0

    def eval(self, x, do_grad=True):
        fx = self.f(x)
        with np.errstate(invalid="ignore", divide="ignore"):
            phi = -np.log(-fx)
        phi[fx >= 0] = np.inf
        self.last_f, self.last_phi = fx, phi
        self.num_feval += 1
        grad_phi = None
        if do_grad:
            gx = self.g(x)
            grad_phi = -gx / fx
            self.last_g, self.last_grad = gx, grad_phi
            self.num_geval += 1
        return phi, grad_phi

    def func(self, x):
        return self.eval(x, do_grad=False)[0]

    def grad(self, x):
        return self.eval(x)[1]

    def is_feasible(self, x):
        return np.all(self.f(x) <= 0)

    def project(self, x):
        return x if self.proj is None else self.proj(x)


def barrier_method(x0, goal_function, reg_dict=None, ineq_dict=None, n_iter=200, t0=1.0, n_biter=10,
                   alpha=0.1, beta_reg=1.0, t_step=0.1, add_stat_cb=None, max_iter=1000,
                   do_alpha_decay=True, min_delta_loss=1e-3, n_steps_without_progress=None,
                   trace=None, on_iter=None):
    """src/barrier_method.py:79-240.  ``trace`` (a list) receives the loss of every early-stop check;
    ``on_iter(i_biter, i_iter, x)`` (test instrumentation, not in the reference) sees every iterate."""
    reg_dict = reg_dict or {}
    ineq_dict = ineq_dict or {}
    objs, bf = {}, {}
    for k, v in ineq_dict.items():                                      # :119-126
        if type(v) is tuple:
            v = BarrierFunction(*v)
        objs[k] = v
        bf[k] = (v.func, v.grad)
    if not all(v.is_feasible(x0) for v in objs.values()):               # :128-152
        for v in objs.values():
            x0 = v.project(x0)
        if not all(v.is_feasible(x0) for v in objs.values()):
            return x0, None
    t, x = t0, x0
    for i_biter in range(n_biter):                                      # :159
        prev_loss, stall = None, 0
        for i_iter in range(n_iter):                                    # :163
            grad = goal_function[1](x)
            reg_grads = [v[1](x) for v in reg_dict.values()]
            bf_grads = [v[1](x) for v in bf.values()]
            for g in reg_grads:
                grad += beta_reg * g                                    # in place, like :168-171
            for g in bf_grads:
                grad += beta_reg * g * t
            if add_stat_cb is not None:                                 # cadence: every inner it (B6)
                add_stat_cb((goal_function[0](x), x.copy(), grad.copy()))
            x = x + (-alpha * grad)                                     # :176,201
            for v in objs.values():                                     # :205-206
                x = v.project(x)
            if on_iter is not None:
                on_iter(i_biter, i_iter, x)
            if n_steps_without_progress is not None:                    # :208-229
                loss = goal_function[0](x)
                for v in reg_dict.values():
                    loss += np.sum(np.ravel(beta_reg * v[0](x)))
                for v in bf.values():
                    loss += np.sum(np.ravel(beta_reg * t * v[0](x)))
                if trace is not None:
                    trace.append(float(loss))
                if prev_loss is not None:
                    stall = 0 if prev_loss - loss >= min_delta_loss else stall + 1
                prev_loss = loss
                if stall >= n_steps_without_progress:
                    break
        t = t * t_step                                                  # :231-236
        if do_alpha_decay:
            alpha = alpha * np.sqrt(t_step)
            n_iter = min(int(n_iter / t_step), max_iter)
    return x, []


class HoughBarrier(BarrierFunction):
    """Sinogram-domain barrier FP(x)[m] >= b (teeth_recon/barrier.py:10-69)."""

    def __init__(self, FP, BP, m, b, r_like):
        self.FP, self.BP, self.m, self.b, self.r = FP, BP, m, b, r_like

        def bc_func(x):
            return b - FP(x)[m]

        def bc_grad(x):
            g = np.zeros_like(r_like).flatten()
            g[m] = -1
            return BP(g)

        def bc_project(x, n_iter=100, alpha=1e-2):                       # :30-45 (B11 kept literally)
            f = FP(x)
            x_new = x.copy()
            for _ in range(n_iter):
                infeas = m & (f < b)
                if not np.any(infeas):
                    break
                f[~infeas] = 0
                f[infeas] -= b + 1e-3
                x_new = x_new - alpha * BP(f)
                f = FP(x_new)
            return x_new

        super(HoughBarrier, self).__init__(bc_func, bc_grad, bc_project)

    def eval(self, x, do_grad=True):                                     # :50-69
        fp = self.FP(x)
        fx = self.b - fp[self.m]
        with np.errstate(invalid="ignore", divide="ignore"):
            phi = -np.log(-fx)
        phi[fx >= 0] = np.inf
        self.last_f, self.last_phi = fx, phi
        self.num_feval += 1
        grad_phi = None
        if do_grad:
            g = np.zeros_like(self.r).flatten()
            g[self.m] = -1 / fx
            grad_phi = -self.BP(g)
            self.last_g, self.last_grad = g, grad_phi
            self.num_geval += 1
        return phi, grad_phi


def ray_transform_from_projection(p, bound, i0):
    """teeth_recon/experiments.py:290-293 (mutates p, like the reference)."""
    m = p <= bound
    p[m] = bound
    return np.log(i0) - np.log(p), m


def barrier_least_squares(FP, BP, v_shape, pr, i0, bound, n_iter=200, n_biter=10, t0=0.2, t_step=0.2,
                          alpha=0.1, max_iter=400, with_sino_barrier=True, do_alpha_decay=False, trace=None,
                          on_iter=None):
    """First run of teeth_recon/experiments.py:651-749 with FP/BP injected (flat in, flat out).

    cost = sum(((FPx-b)[~m]/n_good)^2); grad = 2 BP((FPx-b)[~m]/n_good)  (B8: not the true gradient; kept).
    """
    r, m = ray_transform_from_projection(pr, bound, i0)
    m = m.flatten()
    b = r.flatten()
    bound_log = np.log(i0) - np.log(bound)
    b[m] = bound_log
    bound_log = 0.95 * bound_log
    n_good = m.size - np.sum(m)

    def cost(x):
        d = FP(x) - b
        d[m] = 0
        d /= n_good
        return np.sum(d ** 2)

    def grad(x):
        d = FP(x) - b
        d[m] = 0
        d /= n_good
        return 2.0 * BP(d)

    morezero = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    ineq = {}
    if with_sino_barrier:
        ineq["bound"] = HoughBarrier(FP, BP, m, bound_log, r)
    ineq["morezero"] = morezero
    x0 = 1e-2 + np.zeros(v_shape, dtype="float32").flatten()
    ans, _ = barrier_method(x0, (cost, grad), ineq_dict=ineq, n_iter=n_iter, n_biter=n_biter, t0=t0,
                            t_step=t_step, beta_reg=1.0, alpha=alpha, max_iter=max_iter,
                            do_alpha_decay=do_alpha_decay, trace=trace, on_iter=on_iter)
    return ans.reshape(v_shape)
