"""Full-size checks at BASELINE.json's configurations (C4: 2048^2 x 1800, C5: 8192^2 x 4096) through
size-independent properties (adjointness, linearity, angle-range additivity, zero residual at the ground truth)
plus oracle parity on an angle subsample that the CPU finishes in seconds.  Needs a B200: -m gpu."""
import numpy as np
import pytest

import oracle
from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _discs(n):
    from whitomo_b200 import phantoms
    return phantoms.two_discs(n)


def _parity_tool():
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("parity_fullsize", os.path.join(ROOT, "tools", "parity_fullsize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("config", ["C3", "C4", "C5"])
def test_parity_with_the_float32_reference_at_full_size(config):
    """north_star's bar at the BASELINE configurations: rel. L2 <= 1e-4 between K1/K2 and the float32 restatement of
    ASTRA's running-sum projector -- ALL 720 / 1800 angles of C3 / C4, every 32nd of C5's 4096 (128 angles), one image
    alone and the same image inside a batch of four; FP of the two-disc phantom, BP of its sinogram and of a uniform
    random sinogram.  (The closed-form positions of WT_POS_EXACT are 1.1e-4 / 2.5e-4 / 7.2e-4 away from the reference
    at C4; tools/parity_fullsize.py writes the whole table to profiles/r2_parity_fullsize.json.)"""
    res = _parity_tool().run_config(config, modes=("astra",), with_f64=False)
    for tag in ("single", "batch4"):
        for op in ("fp", "bp_smooth", "bp_random"):
            assert res["astra"][tag][op]["gpu_vs_oracle32"] < 1e-4, (config, tag, op, res["astra"][tag][op])


@pytest.mark.parametrize("n,na_full", [(2048, 1800), (8192, 4096)])
def test_exact_positions_agree_with_the_float64_arbiter(n, na_full):
    """positions="exact": closed-form ray positions.  Same N and D as the full configuration, 8 of its angles
    (axis-aligned, near 45/135 degrees, generic): within 3e-5 (2048^2) / 1e-4 (8192^2) of the float64 closed-form arbiter (what is
    left is the float32 rounding of ASTRA's geometry fields, 2e-4 of a detector pixel per angle at 8192^2, which the
    kernels share with the reference), i.e. as far from the float32 reference as that reference is from exact arithmetic."""
    from whitomo_b200.projector import Projector
    full = np.linspace(0, np.pi, na_full)
    pick = [0, na_full // 4 - 1, na_full // 4, na_full // 2, 3 * na_full // 4, 3 * na_full // 4 + 1, na_full // 7, na_full - 1]
    ang = full[pick]
    d = int(np.sqrt(2) * n + 1)
    g = oracle.Geometry(n, n, d, ang)
    P = Projector(n, n, d, ang, positions="exact")
    x = _discs(n)[0] + 0.5 * _discs(n)[1]
    p = P.fp(torch.from_numpy(x).cuda()).cpu().numpy()
    exact = oracle.fp64(g, x, threads=0)
    tol = 3e-5 if n <= 2048 else 1e-4
    assert rel_l2(p, exact) < tol
    for y in (exact.astype(np.float32), np.random.default_rng(1).random(g.sino_shape).astype(np.float32)):
        v = P.bp(torch.from_numpy(y).cuda()).cpu().numpy()
        assert rel_l2(v, oracle.bp64(g, y, threads=0)) < tol


def test_c4_operator_properties():
    from whitomo_b200.projector import Projector
    n, na = 2048, 1800
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    gen = torch.Generator(device="cuda").manual_seed(12345)
    x = torch.rand((4, n, n), device="cuda", generator=gen)
    y = torch.rand((4,) + P.sino_shape, device="cuda", generator=gen)
    px, bty = P.fp(x), P.bp(y)
    for i in range(4):                                    # <Wx, y> = <x, W^T y>
        lhs = torch.sum(px[i].double() * y[i].double()).item()
        rhs = torch.sum(x[i].double() * bty[i].double()).item()
        assert abs(lhs - rhs) < 2e-6 * abs(lhs)
    lin = P.fp(0.25 * x[0] + 3.0 * x[1])                  # linearity
    assert rel_l2(lin.cpu().numpy(), (0.25 * px[0] + 3.0 * px[1]).cpu().numpy()) < 1e-5
    parts = P.bp(y[0], angle_range=(0, 777)) + P.bp(y[0], angle_range=(777, na))
    assert rel_l2(parts.cpu().numpy(), bty[0].cpu().numpy()) < 1e-5    # single-image calls pair mirrored pixels (rounding-level)
    # mass is conserved exactly where the footprint is a partition of unity (axis-aligned views: half-width 1)
    total = x[0].double().sum().item()
    for a in (0, na - 1):
        assert abs(px[0, a].double().sum().item() - total) < 2e-6 * total


def test_c4_white_beam_zero_residual_and_descent():
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import phantoms
    n, na = 2048, 1800
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    grid, source, ea, pix = phantoms.synth_spectrum(64, n)
    sp = Spectrum(source / np.sum(source * grid[:, 1]), grid[:, 1], ea, pix)
    gt = torch.stack([torch.from_numpy(phantoms.two_discs(n, z=z)) for z in (0.1, 0.6)]).cuda()
    meas, _, _ = P.white_fp(sp, gt, want_qmu=False, want_loss=False)
    assert 0.0 < meas.min().item() and meas.max().item() <= 1.0 + 1e-5     # intensities of a normalised source
    _, qmu, loss = P.white_fp(sp, gt, meas=meas, want_intensity=False)
    assert loss.abs().max().item() < 1e-9 and qmu.abs().max().item() < 1e-6   # residual vanishes at the ground truth
    x = torch.full_like(gt, 0.15)
    _, qmu, l0 = P.white_fp(sp, x, meas=meas, want_intensity=False)
    g = P.bp(qmu.reshape((-1,) + P.sino_shape), scale=-1.0).reshape(x.shape)
    # a small step along -grad must decrease the data term of both slices (gradient sign and scale are right)
    step = 0.05 * l0.min().item() / float((g.double() ** 2).sum().item() / 2)
    _, _, l1 = P.white_fp(sp, x - step * g, meas=meas, want_intensity=False, want_qmu=False)
    assert (l1 < l0).all()


def test_c5_single_slice_adjoint():
    from whitomo_b200.projector import Projector
    n, na = 8192, 4096
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand((n, n), device="cuda", generator=gen)
    y = torch.rand(P.sino_shape, device="cuda", generator=gen)
    lhs = torch.sum(P.fp(x).double() * y.double()).item()
    rhs = torch.sum(x.double() * P.bp(y).double()).item()
    assert abs(lhs - rhs) < 5e-6 * abs(lhs)
    # angle-range split of the oversized slice: 8 partial backprojections add up
    from whitomo_b200.distributed import angle_ranges
    acc = torch.zeros_like(x)
    for r in angle_ranges(na, 8):
        acc += P.bp(y, angle_range=r)
    assert rel_l2(acc.cpu().numpy(), P.bp(y).cpu().numpy()) < 2e-5     # fp32 reassociation of 4096 angle terms
