"""The device soft-inequality least-squares loop (whitomo_b200.solvers.ineq_linear_least_squares) against the
reference's own closures handed to scipy.optimize.fmin_cg (teeth_recon/experiments.py:513-582), both over the CPU oracle
(the loop only needs the Projector interface; tests/fake_projector.py provides it without a GPU)."""
import numpy as np
import pytest

import oracle
from oracle import barrier as ob
from conftest import rel_l2

torch = pytest.importorskip("torch")


def _problem(n=32, na=40):
    g = oracle.Geometry.standard(n, na)
    ph = np.zeros((n, n))
    yy, xx = np.mgrid[0:n, 0:n]
    ph[(xx - 15.5) ** 2 + (yy - 15.5) ** 2 <= 144] = 0.10358164     # misc/calc_recon_areas.py:15
    ph[(xx - 18) ** 2 + (yy - 12) ** 2 <= 9] = 9.21754054
    rng = np.random.RandomState(12345)
    i0, bound, alpha = 1e9, 8.0, 300.0
    pr = rng.poisson(lam=np.exp(np.log(i0) - oracle.fp(g, ph).astype(np.float64)).flatten()).astype('float32').reshape(g.sino_shape)
    return g, ph, pr, i0, bound, alpha


def reference_closures(g, pr, i0, bound, alpha, FP, BP):
    """cost / grad exactly as teeth_recon/experiments.py:531-553 (including the literal `d[1 - m]` index)"""
    r, m = ob.ray_transform_from_projection(pr.copy(), bound, i0)
    m = m.flatten()
    b = r.flatten()
    b[m] = np.log(i0) - np.log(bound)
    n_bad = np.sum(m)
    n_good = m.size - n_bad

    def weigh(x, good_w, bad_w):
        d = FP(x) - b
        d[(d > 0) * m] = 0.0
        d[m] *= alpha
        d[1 - m] /= good_w
        d[m] /= bad_w
        return d
    return (lambda x: np.sum(weigh(x, np.sqrt(n_good), np.sqrt(n_bad)) ** 2)), (lambda x: 2.0 * BP(weigh(x, n_good, n_bad)))


def test_device_loop_reaches_the_stationary_point_scipy_cg_heads_for():
    """scipy's fmin_cg drives the reference's `grad` field towards zero; that field is the gradient of G (docstring of
    ineq_linear_least_squares), NOT of the reference's `cost` (1/n vs 1/sqrt(n) weights), so scipy's line search gives up
    after a step or two on this problem (alpha = 300: gradient norm 1.7e3, cost 1.2e5, for any maxiter).  The device
    loop must reach the stationary point scipy heads for: the reference's own gradient field vanishes there and the
    reference's own cost is far below where scipy stops."""
    import scipy.optimize as opt
    from fake_projector import OracleProjector
    from whitomo_b200 import solvers
    g, ph, pr, i0, bound, alpha = _problem()
    n = g.rows
    FP = lambda x: oracle.fp(g, x.reshape(n, n)).flatten().astype(np.float64)
    cost, grad = reference_closures(g, pr, i0, bound, alpha, FP,
                                    lambda x: oracle.bp(g, x.reshape(g.sino_shape)).flatten().astype(np.float64))
    x_ref = opt.fmin_cg(cost, np.zeros(n * n), grad, maxiter=150, disp=False)
    P = OracleProjector(n, n, g.ndet, np.linspace(0, np.pi, g.nangles))
    x, hist = solvers.ineq_linear_least_squares(P, pr.copy(), i0, bound, alpha, n_iters=150)
    hist = hist.numpy()
    assert np.all(np.diff(hist) <= 1e-9 * hist[0]) and hist[-1] < 1e-4 * hist[0]      # monotone, converging
    xd = x.numpy().astype(np.float64).flatten()
    gn_dev, gn_ref, gn_0 = np.linalg.norm(grad(xd)), np.linalg.norm(grad(x_ref)), np.linalg.norm(grad(np.zeros(n * n)))
    assert gn_dev < 1e-3 * gn_0 and gn_dev <= gn_ref              # the reference's gradient field vanishes ...
    assert cost(xd) < 1e-3 * cost(x_ref)                          # ... and its cost is far below scipy's stopping point
    # the constraints hold: starved rays are at or above log(i0 / bound) (up to the soft penalty), good rays are fitted
    r, m = ob.ray_transform_from_projection(pr.copy(), bound, i0)
    p = FP(xd).reshape(g.sino_shape)
    assert np.all(p[m] > np.log(i0 / bound) - 0.05)
    assert rel_l2(p[~m], r[~m]) < 2e-2
    # one exact line search step from x = 0 equals the minimiser of G along -grad (closed form for a quadratic piece)
    x1, h1 = solvers.ineq_linear_least_squares(P, pr.copy(), i0, bound, alpha, n_iters=1)
    assert h1[1] < h1[0]
