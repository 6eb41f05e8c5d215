"""Parity of the sm_100a kernels (through the C ABI) against the CPU oracle.  Needs a B200: -m gpu.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-4 on sinograms / backprojections,
<= 1e-3 on reconstructions after N iterations.  The kernels evaluate the ray position in closed form
(c0 + l*delta) where ASTRA accumulates a float32 running sum, so agreement is to fp32 rounding, not bitwise.
"""
import numpy as np
import pytest

import oracle
from oracle import white as ow
from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL_OP = 1e-4
TOL_REC = 1e-3


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from whitomo_b200.projector import Projector, Spectrum
    return Projector, Spectrum


def phantom(n, m=None, seed=0):
    m = m or n
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:m]
    x = np.zeros((n, m), np.float32)
    x[(xx - 0.31 * m) ** 2 + (yy - 0.31 * n) ** 2 <= (0.19 * min(n, m)) ** 2] = 1.0
    x[(xx - 0.78 * m) ** 2 + (yy - 0.78 * n) ** 2 <= (0.08 * min(n, m)) ** 2] = 0.5
    x += 0.05 * rng.random((n, m)).astype(np.float32)
    return x


SHAPES = [  # rows, cols, ndet, nangles
    (10, 10, 15, 35), (64, 64, 91, 90), (50, 70, 40, 37), (33, 21, 60, 19), (256, 256, 363, 180),
    (129, 130, 200, 64), (96, 96, 136, 33),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_fp_bp_vs_oracle(eng, shape):
    Projector, _ = eng
    r, c, d, a = shape
    g = oracle.Geometry(r, c, d, np.linspace(0, np.pi, a))
    P = Projector(r, c, d, np.linspace(0, np.pi, a))
    x = phantom(r, c, 1)
    y = np.random.default_rng(2).random((a, d)).astype(np.float32)
    p = P.fp(torch.from_numpy(x).cuda()).cpu().numpy()
    assert p.shape == (a, d) and p.dtype == np.float32
    assert rel_l2(p, oracle.fp(g, x)) < TOL_OP
    assert rel_l2(p, oracle.fp64(g, x)) < TOL_OP
    v = P.bp(torch.from_numpy(y).cuda()).cpu().numpy()
    assert v.shape == (r, c)
    assert rel_l2(v, oracle.bp(g, y)) < TOL_OP
    assert rel_l2(v, oracle.bp64(g, y)) < TOL_OP


def test_special_angles(eng):
    """exact 45/135 degree tie (strict '<' on float32 values), 0, 90, 180 degrees, real-data style 0.5 degree steps"""
    Projector, _ = eng
    ang = np.concatenate([[np.pi / 4, 3 * np.pi / 4, np.pi / 2, 0.0, np.pi], np.deg2rad(np.arange(0, 200, 0.5))[::7]])
    g = oracle.Geometry(48, 48, 69, ang)
    P = Projector(48, 48, 69, ang)
    x = phantom(48, 48, 3)
    y = np.random.default_rng(4).random(g.sino_shape).astype(np.float32)
    assert rel_l2(P.fp(torch.from_numpy(x).cuda()).cpu().numpy(), oracle.fp(g, x)) < TOL_OP
    assert rel_l2(P.bp(torch.from_numpy(y).cuda()).cpu().numpy(), oracle.bp(g, y)) < TOL_OP


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8])
def test_image_batches_cover_all_interleave_widths(eng, n):
    """n images are processed as float4 groups (the last one padded), or as one float2 / float group; every image must equal its solo result"""
    Projector, _ = eng
    r, c, d, a = 40, 56, 75, 24
    g = oracle.Geometry(r, c, d, np.linspace(0, np.pi, a))
    P = Projector(r, c, d, np.linspace(0, np.pi, a))
    xs = np.stack([phantom(r, c, 10 + i) for i in range(n)])
    ys = np.random.default_rng(5).random((n, a, d)).astype(np.float32)
    ps = P.fp(torch.from_numpy(xs).cuda()).cpu().numpy()
    vs = P.bp(torch.from_numpy(ys).cuda()).cpu().numpy()
    for i in range(n):
        assert rel_l2(ps[i], oracle.fp(g, xs[i])) < TOL_OP
        assert rel_l2(vs[i], oracle.bp(g, ys[i])) < TOL_OP
        # batch composition must not change a single bit of an image's result (any batch of >= 3 images) ...
        trio_x = torch.from_numpy(np.stack([xs[i], xs[0], xs[-1]])).cuda()
        trio_y = torch.from_numpy(np.stack([ys[i], ys[0], ys[-1]])).cuda()
        if n >= 3:
            assert np.array_equal(ps[i], P.fp(trio_x)[0].cpu().numpy())
            assert np.array_equal(vs[i], P.bp(trio_y)[0].cpu().numpy())
        # ... while calls with one or two images pair each ray with its point reflection (rounding-level difference)
        assert rel_l2(P.fp(torch.from_numpy(xs[i]).cuda()).cpu().numpy(), P.fp(trio_x)[0].cpu().numpy()) < 1e-5
        assert rel_l2(P.bp(torch.from_numpy(ys[i]).cuda()).cpu().numpy(), P.bp(trio_y)[0].cpu().numpy()) < 1e-5
        assert rel_l2(P.fp(trio_x[:2])[0].cpu().numpy(), P.fp(trio_x)[0].cpu().numpy()) < 1e-5
        assert rel_l2(P.bp(trio_y[:2])[0].cpu().numpy(), P.bp(trio_y)[0].cpu().numpy()) < 1e-5


def test_adjointness_on_device(eng):
    Projector, _ = eng
    P = Projector(200, 200, 283, np.linspace(0, np.pi, 150))
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.random((200, 200)).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.random((150, 283)).astype(np.float32)).cuda()
    lhs = torch.sum(P.fp(x).double() * y.double()).item()
    rhs = torch.sum(x.double() * P.bp(y).double()).item()
    assert abs(lhs - rhs) < 2e-6 * abs(lhs)


def test_bp_angle_ranges_sum_to_full(eng):
    """the angle-range split used for one oversized slice: partial backprojections add up to the full one"""
    Projector, _ = eng
    P = Projector(64, 64, 91, np.linspace(0, np.pi, 50))
    y = torch.from_numpy(np.random.default_rng(7).random((50, 91)).astype(np.float32)).cuda()
    full = P.bp(y)
    parts = sum(P.bp(y, angle_range=(a0, a1)) for a0, a1 in ((0, 13), (13, 14), (14, 37), (37, 50)))
    assert rel_l2(parts.cpu().numpy(), full.cpu().numpy()) < 1e-6
    assert torch.count_nonzero(P.bp(y, angle_range=(5, 5))).item() == 0


def test_g1_and_g2_golden_through_the_engine(eng, golden):
    Projector, _ = eng
    z = golden("g1_teeth")
    pr = z["projections_T"].T
    P = Projector(10, 10, 15, np.linspace(0, np.pi, 35))
    p = P.fp(torch.from_numpy(z["phantom"].astype(np.float32)).cuda()).cpu().numpy()
    empty, starved = pr > 9e8, pr == 0
    assert np.abs(p[empty]).max() < 1e-3 and p[starved].min() > 21.0
    mid = ~empty & ~starved
    assert np.all(np.abs(p[mid] + np.log(pr[mid] / 1e9)) < 5 / np.sqrt(pr[mid]) + 1e-3)

    z = golden("g2_button4")
    shape = tuple(z["shape"])
    c = np.stack([np.unpackbits(z[k])[:shape[0] * shape[1]].reshape(shape).astype(np.float32)
                  for k in ("c1_bits", "c2_bits")])
    P = Projector(shape[0], shape[1], int(np.sqrt(2) * shape[0] + 1), np.linspace(0, np.pi, int(z["n_angles"])))
    grid = z["grid"]
    from whitomo_b200.projector import Spectrum
    src = z["source"] / np.sum(z["source"] * grid[:, 1])
    sp = Spectrum(src, grid[:, 1], z["element_absorptions"], float(z["pixel_size"]))
    inten, _, _ = P.white_fp(sp, torch.from_numpy(c).cuda(), want_qmu=False, want_loss=False)
    rec = P.fbp(torch.log(1.0 / inten[0])).cpu().numpy()
    assert rel_l2(rec, z["fbp"]) < TOL_OP * 2          # the reference's own ASTRA output (src/fbp.nptxt)


def test_sirt_vs_oracle(eng):
    Projector, _ = eng
    n, a = 64, 48
    g = oracle.Geometry.standard(n, a)
    P = Projector(n, n, g.ndet, np.linspace(0, np.pi, a))
    x = phantom(n, n, 8)
    p = oracle.fp(g, x)
    for iters in (1, 30):
        rec = P.sirt(torch.from_numpy(p).cuda(), iters).cpu().numpy()
        assert rel_l2(rec, oracle.sirt(g, p, iters)) < TOL_REC
    two = P.sirt(torch.from_numpy(np.stack([p, 2 * p, 0.5 * p])).cuda(), 5).cpu().numpy()
    ref = oracle.sirt(g, p, 5)
    assert rel_l2(two[1], 2 * ref) < TOL_REC and rel_l2(two[2], 0.5 * ref) < TOL_REC


def _white_setup(eng, z, tag):
    Projector, Spectrum = eng
    n, na = int(z[tag + "_n"]), int(z[tag + "_n_angles"])
    grid, ea, pix = z[tag + "_grid"], z[tag + "_ea"], float(z[tag + "_pixel_size"])
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    src = z[tag + "_source"] / np.sum(z[tag + "_source"] * grid[:, 1])
    return P, Spectrum(src, grid[:, 1], ea, pix)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_white_fp_vs_reference_module_golden(eng, golden, tag):
    """fused FP + exp/spectrum-sum/residual/q*mu vs the outputs of the reference's own white_projection.py"""
    z = golden("white_small")
    P, sp = _white_setup(eng, z, tag)
    gt = torch.from_numpy(z[tag + "_gt"].astype(np.float32)).cuda()
    c = torch.from_numpy(z[tag + "_c"].astype(np.float32)).cuda()
    meas, _, _ = P.white_fp(sp, gt, want_qmu=False, want_loss=False)
    assert rel_l2(meas.cpu().numpy().ravel(), z[tag + "_sinogram"]) < TOL_OP
    inten, qmu, loss = P.white_fp(sp, c, meas=meas)
    assert rel_l2(inten.cpu().numpy().ravel(), z[tag + "_I"]) < TOL_OP
    assert abs(loss.item() - float(z[tag + "_func"])) < 1e-4 * float(z[tag + "_func"])
    q = z[tag + "_sinogram"] - z[tag + "_I"]
    # measured 5e-7 (q mu) and 2e-6 (gradient): profiles/r2_test_margins.json
    assert rel_l2(qmu[0].cpu().numpy().reshape(2, -1), q * z[tag + "_mu"]) < TOL_OP
    grad = -P.bp(qmu[0]).cpu().numpy()
    assert rel_l2(grad, z[tag + "_grad"]) < TOL_OP


def test_white_fp_stack_and_odd_k(eng, golden):
    """slices x elements interleave (K=2 -> 2 slices per float4 group) and the generic-K path agree with the oracle"""
    Projector, Spectrum = eng
    rng = np.random.default_rng(9)
    n, na, E = 40, 30, 7
    for K in (1, 2, 3):
        grid = np.stack([np.linspace(1, 9, E), rng.uniform(0.5, 1.5, E)]).T
        source = rng.uniform(0.2, 2.0, E)
        ea = rng.uniform(200, 4000, (K, E))
        pix = 5e-6
        wo = ow.WhiteOracle(source, pix, ea, grid, (n, n), n_angles=na)
        P = Projector(n, n, wo.geom.ndet, np.linspace(0, np.pi, na))
        sp = Spectrum(wo.source, grid[:, 1], ea, pix)
        S = 3
        conc = rng.uniform(0.0, 0.3, (S, K, n, n)).astype(np.float32)
        meas = rng.uniform(0.2, 1.0, (S,) + wo.proj_shape).astype(np.float32)
        inten, qmu, loss = P.white_fp(sp, torch.from_numpy(conc).cuda(), meas=torch.from_numpy(meas).cuda())
        for s in range(S):
            I, _, ex, _ = wo.FP_white(conc[s])
            q = meas[s].ravel() - I
            assert rel_l2(inten[s].cpu().numpy().ravel(), I) < TOL_OP
            assert rel_l2(qmu[s].cpu().numpy().reshape(K, -1), q * wo.calc_mu(ex)) < TOL_OP
            assert abs(loss[s].item() - 0.5 * np.sum(q ** 2)) < 1e-4 * 0.5 * np.sum(q ** 2)


def test_pair_norm2_vs_numpy(eng):
    """wt_pair_norm2: ||c0 c1|| of the loss Iteration reports (src/white_rec.py:236-237), ragged pixel counts included"""
    Projector, _ = eng
    rng = np.random.default_rng(21)
    for S, r, c in ((1, 32, 32), (3, 70, 90), (2, 257, 129)):
        P = Projector(r, c, int(np.hypot(r, c)) + 2, np.linspace(0, np.pi, 4))
        x = rng.uniform(0.0, 1.0, (S, 2, r, c)).astype(np.float32)
        part = P.pair_norm2(torch.from_numpy(x).cuda())
        assert part.shape == (S, 32) and part.dtype == torch.float64
        got = torch.sqrt(part.sum(-1)).cpu().numpy()
        ref = [np.linalg.norm((x[s, 0] * x[s, 1]).astype(np.float64)) for s in range(S)]
        assert np.allclose(got, ref, rtol=1e-6)
        again = P.pair_norm2(torch.from_numpy(x).cuda())
        assert torch.equal(part, again)          # fixed summation order


def test_barrier_update_vs_numpy(eng):
    Projector, _ = eng
    P = Projector(32, 32, 46, np.linspace(0, np.pi, 8))
    rng = np.random.default_rng(10)
    for K, reg_w, lower_only in ((2, 0.05, False), (2, 0.0, False), (1, 0.0, True), (3, 0.0, False)):
        S = 3
        x = rng.uniform(0.02, 0.98, (S, K, 32, 32)).astype(np.float32)
        g = rng.normal(0, 0.5, x.shape).astype(np.float32)
        alpha, beta, t = 0.3, 1.0, 0.05
        grad = g.astype(np.float64)
        if reg_w:
            grad = grad + beta * reg_w * np.stack([ow.calc_uneq_reg_grad(xx.astype(np.float64)) for xx in x])
        grad = grad + beta * t * (-1.0 / x)
        if not lower_only:
            grad = grad + beta * t * (1.0 / (1.0 - x))
        xn = np.clip(x - alpha * grad, 0, None if lower_only else 1)
        xt = torch.from_numpy(x.copy()).cuda()
        pen = P.barrier_update(xt, torch.from_numpy(g).cuda(), alpha, beta, t, reg_w=reg_w, lower_only=lower_only,
                               want_penalty=True)
        assert rel_l2(xt.cpu().numpy(), xn) < 1e-6
        xn32 = xt.cpu().numpy().astype(np.float64)
        for s in range(S):
            with np.errstate(divide="ignore"):
                ph = -np.log(xn32[s]) if lower_only else -np.log(xn32[s]) - np.log(1 - xn32[s])
            ref = beta * t * ph.sum() + (beta * reg_w * np.abs(xn32[s, 0] * xn32[s, 1]).sum() if reg_w else 0.0)
            if np.isfinite(ref):
                assert abs(pen[s].item() - ref) < 1e-4 * abs(ref)
            else:
                assert not np.isfinite(pen[s].item())


def test_errors_are_loud(eng):
    Projector, _ = eng
    from whitomo_b200._capi import EngineError
    with pytest.raises(EngineError):
        Projector(16, 16, 23, np.linspace(0, np.pi, 4), det_width=0.5)
    P = Projector(16, 16, 23, np.linspace(0, np.pi, 4))
    with pytest.raises(ValueError):
        P.fp(torch.zeros(15, 16, device="cuda"))
    with pytest.raises(EngineError):
        P.bp(torch.zeros(4, 23, device="cuda"), angle_range=(3, 9))


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 5, 3, 2), (2, 2, 4, 3), (3, 1, 7, 5), (7, 300, 9, 4), (300, 7, 500, 3)])
def test_degenerate_and_ragged_geometries(eng, shape):
    """1-pixel volumes, 1-detector sinograms, a single view, extreme aspect ratios, detector far wider/narrower
    than the object"""
    Projector, _ = eng
    r, c, d, a = shape
    ang = np.linspace(0.3, 2.9, a)
    g = oracle.Geometry(r, c, d, ang)
    P = Projector(r, c, d, ang)
    rng = np.random.default_rng(r * 1000 + c)
    x = rng.random((3, r, c)).astype(np.float32)
    y = rng.random((3, a, d)).astype(np.float32)
    p = P.fp(torch.from_numpy(x).cuda()).cpu().numpy()
    v = P.bp(torch.from_numpy(y).cuda()).cpu().numpy()
    for i in range(3):
        ref_p, ref_v = oracle.fp(g, x[i]), oracle.bp(g, y[i])
        assert np.abs(p[i] - ref_p).max() <= 1e-4 * max(1.0, np.abs(ref_p).max())
        assert np.abs(v[i] - ref_v).max() <= 1e-4 * max(1.0, np.abs(ref_v).max())


def test_empty_batch_and_zero_iterations(eng):
    Projector, _ = eng
    P = Projector(16, 16, 23, np.linspace(0, np.pi, 6))
    assert P.fp(torch.zeros((0, 16, 16), device="cuda")).shape == (0, 6, 23)
    assert P.bp(torch.zeros((0, 6, 23), device="cuda")).shape == (0, 16, 16)
    z = P.sirt(torch.rand(6, 23, device="cuda"), 0)
    assert z.shape == (16, 16) and torch.count_nonzero(z).item() == 0
    assert torch.count_nonzero(P.fp(torch.zeros(16, 16, device="cuda"))).item() == 0


def test_one_projector_on_two_streams(eng):
    """calls issued on different CUDA streams may overlap on the device: each stream gets its own scratch, so the
    results equal those of the same calls issued one after the other"""
    Projector, _ = eng
    n, a = 192, 120
    d = int(np.sqrt(2) * n + 1)
    P = Projector(n, n, d, np.linspace(0, np.pi, a))
    x1 = torch.from_numpy(np.stack([phantom(n, seed=s) for s in range(4)])).cuda()
    x2 = torch.from_numpy(np.stack([phantom(n, seed=10 + s) for s in range(4)])).cuda()
    ref = [P.fp(x1).clone(), P.fp(x2).clone(), P.bp(P.fp(x1)).clone(), P.bp(P.fp(x2)).clone()]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            p1 = P.fp(x1)
            v1 = P.bp(p1)
        with torch.cuda.stream(s2):
            p2 = P.fp(x2)
            v2 = P.bp(p2)
        torch.cuda.synchronize()
        for got, want in zip((p1, p2, v1, v2), ref):
            assert torch.equal(got, want)
    assert len(P._ws) >= 2


@pytest.mark.parametrize("det_width", [1.0, 1.25, 2.0])
def test_wide_detector_pixels_and_unordered_angles(eng, det_width):
    """detector pixels wider than a volume pixel (create_proj_geom's det_width argument) and an angle list that is
    unsorted, repeats a view and leaves [0, pi)"""
    Projector, _ = eng
    r, c, d = 70, 90, 101
    ang = np.array([2.1, -0.4, 0.0, 3.9, 0.77, 0.77, 1.5707963, 6.0, 1.2, 2.8, -2.2])
    g = oracle.Geometry(r, c, d, ang, det_width=det_width)
    P = Projector(r, c, d, ang, det_width=det_width)
    x = np.stack([phantom(r, c, s) for s in range(5)])
    y = np.random.default_rng(8).random((5,) + g.sino_shape).astype(np.float32)
    p = P.fp(torch.from_numpy(x).cuda()).cpu().numpy()
    v = P.bp(torch.from_numpy(y).cuda()).cpu().numpy()
    for i in range(5):
        assert rel_l2(p[i], oracle.fp(g, x[i])) < TOL_OP
        assert rel_l2(v[i], oracle.bp(g, y[i])) < TOL_OP
    assert np.array_equal(p[:, 4], p[:, 5])                # the repeated view


def test_wide_detector_pixels_on_a_volume_wider_than_the_staging_window(eng):
    """det_width = 1.5 on 300 x 260 (lines longer than K1's 96-position window, so the window really slides): parity;
    det_width = 2 near 45 degrees spreads one angle's 32 detectors x 16 lines over more than the window holds:
    the geometry is refused (WT_ERR_UNSUPPORTED) instead of being projected wrongly"""
    Projector, _ = eng
    from whitomo_b200._capi import EngineError
    r, c, d = 300, 260, 301
    ang = np.linspace(0, np.pi, 37)
    g = oracle.Geometry(r, c, d, ang, det_width=1.5)
    P = Projector(r, c, d, ang, det_width=1.5)
    x = np.stack([phantom(r, c, s) for s in range(3)])
    y = np.random.default_rng(9).random((3,) + g.sino_shape).astype(np.float32)
    p = P.fp(torch.from_numpy(x).cuda()).cpu().numpy()
    v = P.bp(torch.from_numpy(y).cuda()).cpu().numpy()
    for i in range(3):
        assert rel_l2(p[i], oracle.fp(g, x[i])) < TOL_OP
        assert rel_l2(v[i], oracle.bp(g, y[i])) < TOL_OP
    with pytest.raises(EngineError, match="too wide"):
        Projector(r, c, d, ang, det_width=2.0)


def test_repeated_calls_are_bit_identical(eng):
    """the TMA / mbarrier pipelines hand stages between a producer warp and 8 consumer warps with no other
    synchronisation; a hand-off race would show up as run-to-run differences"""
    Projector, _ = eng
    n, a = 384, 300
    d = int(np.sqrt(2) * n + 1)
    P = Projector(n, n, d, np.linspace(0, np.pi, a))
    g = torch.Generator(device="cuda").manual_seed(7)
    for nimg in (1, 2, 6):
        x = torch.rand(nimg, n, n, device="cuda", generator=g)
        y = torch.rand(nimg, a, d, device="cuda", generator=g)
        p0, v0 = P.fp(x).clone(), P.bp(y).clone()
        for _ in range(10):
            assert torch.equal(P.fp(x), p0) and torch.equal(P.bp(y), v0)
