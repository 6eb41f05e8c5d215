"""Pins the CPU oracle (oracle/) on the reference's own artifacts and on outputs of the reference's own
Python modules (tests/golden/*.npz, generator: tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import white as ow, barrier as ob
from conftest import rel_l2


def test_g1_teeth_footprint(golden):
    """teeth_recon/phantom.txt -> teeth_recon/projjections.txt (ASTRA FP + Poisson, i0=1e9): every one of the
    525 rays must be on the right side of empty / photon-starved, and the 6 mid-range rays must agree within
    the fixture's Poisson noise.  Pins orientation (row 0 top, detector/angle order) and the 2-tap footprint."""
    z = golden("g1_teeth")
    pr = z["projections_T"].T
    g = oracle.Geometry.standard(10, 35)
    assert g.sino_shape == pr.shape == (35, 15)
    p = oracle.fp(g, z["phantom"])
    empty, starved = pr > 9e8, pr == 0
    assert empty.sum() == 228 and starved.sum() == 291
    assert np.abs(p[empty]).max() < 1e-3
    assert p[starved].min() > 21.0
    mid = ~empty & ~starved
    assert mid.sum() == 6
    meas = -np.log(pr[mid] / 1e9)
    sigma = 1.0 / np.sqrt(pr[mid])                     # relative Poisson sd -> absolute sd of the log
    assert np.all(np.abs(p[mid] - meas) < 5 * sigma + 1e-3)


def _g2_chain(z, fp):
    shape = tuple(z["shape"])
    c = [np.unpackbits(z[k])[:shape[0] * shape[1]].reshape(shape).astype(np.float64) for k in ("c1_bits", "c2_bits")]
    g = oracle.Geometry.standard(shape[0], int(z["n_angles"]))
    grid, ea, pix = z["grid"], z["element_absorptions"], float(z["pixel_size"])
    src = z["source"] / np.sum(z["source"] * grid[:, 1])
    P = np.array([fp(g, cc) for cc in c]).astype(np.float64).reshape(2, -1)
    I = np.sum(src * np.exp(-pix * np.einsum("kr,ke->re", P, ea)) * grid[:, 1], axis=1)
    return g, np.log(1.0 / I).reshape(g.sino_shape)


def test_g2_fbp_chain(golden):
    """src/fbp.nptxt = cpu_fbp(log(1/FP_white(button_4 gt))) (src/barrier_method_gogo_additional_plots.py:93-97):
    pins FP, the white-beam pipeline, BP and the FBP filter/scale in one go."""
    z = golden("g2_button4")
    g, log_sino = _g2_chain(z, oracle.fp)
    rec = oracle.fbp(g, log_sino)
    assert rel_l2(rec, z["fbp"]) < 1e-4
    # discrimination: a vertical flip or swapped elements must be rejected by a wide margin
    assert rel_l2(rec[::-1], z["fbp"]) > 0.5


def test_adjoint_and_gather_form():
    rng = np.random.default_rng(12345)
    for (r, c, d, a) in ((64, 64, 91, 90), (50, 70, 40, 37), (33, 21, 60, 19)):
        g = oracle.Geometry(r, c, d, np.linspace(0, np.pi, a))
        x, y = rng.random((r, c)), rng.random((a, d))
        lhs = np.vdot(oracle.fp(g, x).astype(np.float64), y)
        rhs = np.vdot(x, oracle.bp(g, y).astype(np.float64))
        assert abs(lhs - rhs) < 1e-5 * abs(lhs)
        assert rel_l2(oracle.bp(g, y), oracle.bp64(g, y)) < 2e-5      # scatter == pixel-driven gather (A.3)
        assert rel_l2(oracle.fp(g, x), oracle.fp64(g, x)) < 2e-5
        assert np.array_equal(oracle.fp(g, x), oracle.fp(g, x, threads=3))
        assert rel_l2(oracle.bp(g, y, threads=3), oracle.bp(g, y)) < 1e-5


def test_45_degree_tie_and_odd_angles():
    g = oracle.Geometry(32, 32, 47, np.array([np.pi / 4, 3 * np.pi / 4, np.pi / 2, 0.0, np.pi, 2.5 * np.pi / 180 * 3]))
    x = np.random.default_rng(1).random((32, 32))
    assert rel_l2(oracle.fp(g, x), oracle.fp64(g, x)) < 2e-5
    y = np.random.default_rng(2).random(g.sino_shape)
    assert rel_l2(oracle.bp(g, y), oracle.bp64(g, y)) < 2e-5


def test_sirt_reduces_residual_unpinned():
    """SIRT has no ASTRA artifact in the reference (parity unpinned); check the A.4 recurrence is a contraction."""
    g = oracle.Geometry.standard(32, 30)
    x = np.zeros((32, 32), np.float32)
    x[8:20, 10:25] = 1.0
    p = oracle.fp(g, x)
    r10, r50 = oracle.sirt(g, p, 10), oracle.sirt(g, p, 50)
    assert np.linalg.norm(oracle.fp(g, r50) - p) < np.linalg.norm(oracle.fp(g, r10) - p) < np.linalg.norm(p)
    assert rel_l2(r50, x) < 0.25


def _white(z, tag):
    n, na = int(z[tag + "_n"]), int(z[tag + "_n_angles"])
    wo = ow.WhiteOracle(z[tag + "_source"], float(z[tag + "_pixel_size"]), z[tag + "_ea"], z[tag + "_grid"],
                        (n, n), n_angles=na)
    wo.calc_sinogram_with_gt(z[tag + "_gt"])
    return wo


@pytest.mark.parametrize("tag", ["a", "b"])
def test_white_restatement_vs_reference_module(golden, tag):
    z = golden("white_small")
    wo = _white(z, tag)
    c = z[tag + "_c"]
    assert rel_l2(wo.sinogram, z[tag + "_sinogram"]) < 1e-12
    assert abs(wo.func(c) - float(z[tag + "_func"])) < 1e-12 * abs(float(z[tag + "_func"]))
    assert rel_l2(wo.grad(c), z[tag + "_grad"]) < 1e-12
    I, _, exp, _ = wo.FP_white(c)
    assert rel_l2(I, z[tag + "_I"]) < 1e-12
    assert rel_l2(wo.calc_mu(exp), z[tag + "_mu"]) < 1e-12


def test_white_gradient_finite_differences(golden):
    z = golden("white_small")
    wo = _white(z, "b")
    c = z["b_c"].copy()
    g = wo.grad(c)
    rng = np.random.default_rng(3)
    for _ in range(3):
        k, i, j = rng.integers(2), rng.integers(4, 20), rng.integers(4, 20)
        h = 1e-3
        cp, cm = c.copy(), c.copy()
        cp[k, i, j] += h
        cm[k, i, j] -= h
        fd = (wo.func(cp) - wo.func(cm)) / (2 * h)
        assert abs(fd - g[k, i, j]) < 2e-3 * abs(fd) + 1e-9


def test_barrier_white_vs_reference_module(golden):
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    wo = ow.WhiteOracle(z["source"], float(z["pixel_size"]), z["ea"], z["grid"], (n, n), n_angles=na)
    wo.calc_sinogram_with_gt(z["gt"])
    ir = float(z["ir_beta"])
    reg = (lambda x: ir * np.abs(x[0] * x[1]), lambda x: ir * ow.calc_uneq_reg_grad(x))
    nonneg = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    le_one = (lambda x: x - 1, lambda x: np.ones_like(x), lambda x: np.clip(x, x.min(), 1))
    p = {k[2:]: z[k].item() for k in z.files if k.startswith("p_")}
    ans, _ = ob.barrier_method(z["x0"].copy(), (wo.func, wo.grad), reg_dict={"r": reg},
                               ineq_dict={"a": nonneg, "b": le_one}, **p)
    assert np.isfinite(z["ans"]).all()
    assert rel_l2(ans, z["ans"]) < 1e-12


def test_barrier_mar_vs_reference_function(golden):
    z = golden("barrier_mar")
    n, na = int(z["n"]), int(z["n_angles"])
    g = oracle.Geometry.standard(n, na)
    FP = lambda x: oracle.fp(g, x.reshape(n, n)).flatten()
    BP = lambda x: oracle.bp(g, x.reshape(g.sino_shape)).flatten()
    nb = int(z["n_biter"])
    ans = ob.barrier_least_squares(FP, BP, (n, n), z["projections"].copy(), float(z["i0"]), float(z["bound"]),
                                   n_biter=nb)
    assert np.isfinite(z["ans"]).all()
    assert rel_l2(ans, z["ans"]) < 1e-6
    ans2 = ob.barrier_least_squares(FP, BP, (n, n), z["projections"].copy(), float(z["i0"]), float(z["bound"]),
                                    n_iter=100, n_biter=nb, t0=0.1, t_step=0.2, alpha=0.1, max_iter=1000,
                                    with_sino_barrier=False, do_alpha_decay=True)
    assert rel_l2(ans2, z["ans_no_ineq"]) < 1e-6
