"""The reference-runnable configurations of BASELINE.json at their real sizes (C1, C2, C3): parity against the
oracle where the CPU finishes in tens of seconds, solver sanity where it does not.  Needs a B200: -m gpu."""
import numpy as np
import pytest

import oracle
from oracle import white as ow
from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_c1_mono_fp_bp_and_100_sirt_iterations(golden):
    """configs[0]: 256x256 phantom, 180 angles, monochromatic FP/BP + 100 SIRT iterations via astra_proxy"""
    from whitomo_b200 import astra_proxy as ap
    ph = golden("phantom_line_256")["phantom"].astype(np.float64)
    n, na = 256, 180
    pg = ap.astra.create_proj_geom('parallel', 1.0, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    vg = ap.astra.create_vol_geom(n, n)
    pid = ap.astra.create_projector('linear', pg, vg)
    g = oracle.Geometry.standard(n, na)
    p = ap.gpu_fp(pg, vg, ph, pid)
    assert rel_l2(p, oracle.fp(g, ph)) < 1e-4
    assert rel_l2(ap.gpu_bp(pg, vg, p, pid), oracle.bp(g, p)) < 1e-4
    rec = ap.cpu_sirt(pg, vg, pid, p, n_iters=100)
    assert rel_l2(rec, oracle.sirt(g, p, 100)) < 1e-3
    assert rel_l2(rec, ph) < 0.3


def test_c2_white_rec_512(golden):
    """configs[1]: 512x512, 360 angles, 64-bin spectrum, white_rec gradient reconstruction (200 iterations on the
    device; the first 4 iterations are compared with the oracle's restatement of Iteration)"""
    from whitomo_b200 import phantoms, solvers
    from whitomo_b200.projector import Projector, Spectrum
    n, na = 512, 360
    grid, source, ea, pix = phantoms.synth_spectrum(64, n)
    wo = ow.WhiteOracle(source, pix, ea, grid, (n, n), n_angles=na)
    wo.FP = lambda x: oracle.fp(wo.geom, x, threads=0)          # all host threads: same arithmetic per ray
    gt = phantoms.two_discs(n).astype(np.float64)
    wo.calc_sinogram_with_gt(gt)
    P = Projector(n, n, wo.geom.ndet, np.linspace(0, np.pi, na))
    sp = Spectrum(wo.source, grid[:, 1], ea, pix)
    meas = P.white_fp(sp, torch.from_numpy(gt.astype(np.float32)).cuda(), want_qmu=False, want_loss=False)[0]
    assert rel_l2(meas.cpu().numpy().ravel(), wo.sinogram) < 1e-4
    rng = np.random.default_rng(12345)
    c0 = (0.1 + rng.uniform(-0.01, 0.01, gt.shape)).astype(np.float32)
    c = c0.astype(np.float64)
    for it in range(4):
        c, _ = ow.iteration(wo, c, 0.7, 1e-2, clip_concentrations=True, dissimilarity_constraint=True,
                            triv_conc_constraint=True, iter_num=it)
    cd, _ = solvers.white_rec_stack(P, sp, meas, torch.from_numpy(c0).cuda(), n_iters=4)
    assert rel_l2(cd[0].cpu().numpy(), c) < 1e-3
    cd, losses = solvers.white_rec_stack(P, sp, meas, torch.from_numpy(c0).cuda(), n_iters=200)
    l = [x.item() for x in losses]
    assert np.isfinite(l).all() and l[-1] < 0.5 * l[0]
    assert 0.0 <= cd.min().item() and cd.max().item() <= 1.0


def test_c3_metal_artifact_reduction_1024():
    """configs[2]: 1024x1024 teeth-like phantom with a metal insert, 720 angles, log-barrier MAR"""
    from whitomo_b200 import phantoms, solvers, synthesis
    from whitomo_b200.projector import Projector
    n, na = 1024, 720
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    # the pinned levels are mu*pixel at the reference's 256^2 sampling; 4x finer pixels -> 1/4 per pixel
    ph = torch.from_numpy(phantoms.tooth_phantom(n) * (256.0 / n)).cuda()
    gen = torch.Generator(device="cuda").manual_seed(12345)
    i0, bound = 1e9, 8.0
    counts = synthesis.create_projection_with_poisson_noise(i0, P, ph, generator=gen)
    starved = (counts <= bound).float().mean().item()
    assert 0.0 < starved < 0.2                                   # the gold insert starves a fraction of the rays
    # The reference hard-codes its step sizes for its own 64x64 / 90-angle run (teeth_recon/experiments.py:887-889,
    # 741-749; projection step 1e-2 at teeth_recon/barrier.py:30).  Neither scales with N*A: at 1024^2 the feasibility
    # projection overshoots by ~40x and alpha = 0.1 walks pixels across x = 0, after which the barrier term is -1/0
    # (the same failure the reference's own schedule reaches at 32^2 after 8 outer iterations, see make_golden.py).
    # The scenario is therefore run with a 30x smaller step and two barrier levels; what is checked is what the
    # method promises: a finite, non-negative image that satisfies the sinogram-domain barrier, with the data term
    # going down at every barrier level.
    costs = []
    x = solvers.barrier_least_squares(P, counts, i0, bound, n_iter=200, n_biter=2, alpha=0.003,
                                      on_outer=lambda d: costs.append(d["stats"][0, 0].item()))
    assert x is not None and torch.isfinite(x).all() and x.min().item() >= 0.0
    assert len(costs) == 2 and costs[1] < costs[0]
    m = counts <= bound
    assert (P.fp(x)[m] >= 0.95 * np.log(i0 / bound) - 1e-3).all()
