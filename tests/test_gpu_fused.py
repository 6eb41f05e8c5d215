"""The fused iteration (wt_white_barrier_step: K2 with the update in its epilogue -> K1 writing q*mu packed -> one
reduction) and the fused SIRT loop against the separate entry points they replace.  Needs a B200: -m gpu."""
import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _problem(n, na, S, K=2, E=16):
    from whitomo_b200 import phantoms
    from whitomo_b200.projector import Projector, Spectrum
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    grid, source, ea, pix = phantoms.synth_spectrum(E, n)
    sp = Spectrum(source / np.sum(source * grid[:, 1]), grid[:, 1], ea[:K], pix)
    gt = torch.from_numpy(np.stack([phantoms.two_discs(n, z=s / float(S)) for s in range(S)])[:, :K].copy()).cuda()
    meas, _, _ = P.white_fp(sp, gt, want_qmu=False, want_loss=False)
    gen = torch.Generator(device="cuda").manual_seed(3)
    x0 = (0.3 + 0.2 * torch.rand(gt.shape, device="cuda", generator=gen)).contiguous()
    return P, sp, meas, x0


@pytest.mark.parametrize("n,na,S,K", [(96, 60, 3, 2), (160, 90, 1, 2), (130, 75, 5, 1), (64, 40, 2, 2)])
def test_fused_step_equals_the_separate_entry_points(n, na, S, K):
    """x after every step bit-identical (the same upd_pixel on the same gradient bits), f(x) and the penalty to rounding
    of their reductions"""
    P, sp, meas, x0 = _problem(n, na, S, K)
    assert P.step_supported(sp, S)
    xa, xb = x0.clone(), x0.clone()
    state = P.new_step_state(sp, S)
    la, _ = P.white_barrier_step(sp, xa, meas, state, init=True)
    _, qmu, lb = P.white_fp(sp, xb, meas=meas, want_intensity=False)
    assert torch.equal(xa, x0) and torch.allclose(la, lb, rtol=1e-12)
    active = torch.ones(S, dtype=torch.bool, device="cuda")
    if S > 2:
        active[1] = False                                   # an early-stopped slice must stay untouched
    kw = dict(alpha=0.05, beta_reg=1.0, t=0.05, reg_w=0.05 if K == 2 else 0.0)
    for it in range(4):
        la, pa = P.white_barrier_step(sp, xa, meas, state, active=active, want_penalty=True, **kw)
        g = P.bp(qmu.reshape((S * K,) + P.sino_shape), scale=-1.0)
        pb = P.barrier_update(xb, g, kw["alpha"], kw["beta_reg"], kw["t"], reg_w=kw["reg_w"], active=active, want_penalty=True)
        _, qmu, lb = P.white_fp(sp, xb, meas=meas, want_intensity=False)
        assert torch.equal(xa, xb), it
        assert torch.allclose(la, lb, rtol=1e-12) and torch.allclose(pa, pb, rtol=1e-5)
    if S > 2:
        assert torch.equal(xa[1], x0[1])


def test_fused_white_rec_and_barrier_loops_equal_the_unfused_ones():
    from whitomo_b200 import solvers
    P, sp, meas, x0 = _problem(96, 60, 3)
    ca, la = solvers.white_rec_stack(P, sp, meas, x0, n_iters=9, fused=True)
    cb, lb = solvers.white_rec_stack(P, sp, meas, x0, n_iters=9, fused=False)
    assert torch.equal(ca, cb)
    assert np.allclose(torch.stack(la).cpu().numpy(), torch.stack(lb).cpu().numpy(), rtol=1e-10)
    p = dict(n_iter=6, n_biter=3, t0=0.1, t_step=0.5, alpha=0.05, ir_beta=0.05, max_iter=20, n_steps_without_progress=2,
             min_delta_loss=5.0)                              # a loose progress test: slices stop early and get compacted
    xa, ia = solvers.white_barrier_stack(P, sp, meas, x0, fused=True, **p)
    xb, ib = solvers.white_barrier_stack(P, sp, meas, x0, fused=False, **p)
    assert torch.equal(ia["slice_iters"], ib["slice_iters"]) and ia["iters"] == ib["iters"]
    assert torch.equal(xa, xb)
    assert torch.allclose(ia["loss"], ib["loss"], rtol=1e-10)


def test_sirt_fused_loop_vs_oracle_and_workspace_canary():
    """wt_sirt with exactly wt_sirt_workspace_bytes of scratch followed by a canary: the two-launch loop (K1 residual
    epilogue -> packed sinogram, K2 update epilogue -> volume + packed copies) matches the oracle and stays inside its
    workspace (the round-1 loop carved 0.85 MB past it for single small images)"""
    import ctypes
    import oracle
    from whitomo_b200 import phantoms, _capi
    from whitomo_b200.projector import Projector
    for n, na, nimg in ((256, 180, 1), (96, 50, 3), (64, 33, 2)):
        g = oracle.Geometry.standard(n, na)
        P = Projector(n, n, g.ndet, np.linspace(0, np.pi, na))
        x = np.stack([phantoms.two_discs(n, z=0.2 * i)[0] for i in range(nimg)]).astype(np.float32)
        p = np.stack([oracle.fp(g, x[i]) for i in range(nimg)])
        lib = _capi.load()
        nbytes = int(lib.wt_sirt_workspace_bytes(P.handle, nimg))
        ws = torch.full((nbytes + 4096,), 0x5A, dtype=torch.uint8, device="cuda")
        out = torch.empty((nimg, n, n), dtype=torch.float32, device="cuda")
        pd = torch.from_numpy(p).cuda()
        _capi.check(lib.wt_sirt(P.handle, ctypes.c_void_p(pd.data_ptr()), ctypes.c_void_p(out.data_ptr()), nimg, 12,
                                ctypes.c_void_p(ws.data_ptr()), nbytes, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert bool((ws[nbytes:] == 0x5A).all()), (n, nimg)
        for i in range(nimg):
            assert rel_l2(out[i].cpu().numpy(), oracle.sirt(g, p[i], 12)) < 1e-3
        assert lib.wt_sirt(P.handle, ctypes.c_void_p(pd.data_ptr()), ctypes.c_void_p(out.data_ptr()), nimg, 1,
                           ctypes.c_void_p(ws.data_ptr()), nbytes - 4096, None) == -3          # WT_ERR_WORKSPACE
