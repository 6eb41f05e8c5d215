"""Parity of the host-side mirrors and the fused device solver loops against the oracle / the goldens produced by
the reference's own Python modules.  Needs a B200: -m gpu.  Tolerance on reconstructions: rel. L2 <= 1e-3."""
import numpy as np
import pytest

import oracle
from oracle import white as ow, barrier as ob
from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL_OP, TOL_REC = 1e-4, 1e-3


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_astra_proxy_numpy_api(golden):
    """the reference's call forms: gpu_fp/gpu_bp/cpu_sirt/cpu_fbp with numpy in/out, and the verbatim ASTRA object
    manager sequence of src/astra_proxy.py:35-49 through the shim"""
    from whitomo_b200 import astra_proxy as ap
    astra = ap.astra
    n, na = 64, 48
    pg = astra.create_proj_geom('parallel', 1.0, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    vg = astra.create_vol_geom(n, n)
    pid = astra.create_projector('linear', pg, vg)
    g = oracle.Geometry.standard(n, na)
    x = np.random.default_rng(0).random((n, n))            # float64 in, like the reference
    p = ap.gpu_fp(pg, vg, x, pid)
    assert isinstance(p, np.ndarray) and p.dtype == np.float32 and p.shape == (na, g.ndet)
    assert rel_l2(p, oracle.fp(g, x)) < TOL_OP
    v = ap.gpu_bp(pg, vg, p, pid, supersampling=3)
    assert v.dtype == np.float32 and rel_l2(v, oracle.bp(g, p)) < TOL_OP
    assert rel_l2(ap.cpu_sirt(pg, vg, pid, p, n_iters=20), oracle.sirt(g, p, 20)) < TOL_REC
    assert rel_l2(ap.cpu_fbp(pg, vg, pid, p, n_iters=7), oracle.fbp(g, p)) < TOL_REC
    with pytest.raises(ValueError):
        ap.gpu_fp(pg, vg, np.zeros((n, n + 1)), pid)
    # object-manager form (teeth_recon/experiments.py:103-192 carries this copy of astra_proxy)
    apx = ap.AstraProxy(2)
    v_id = apx.data_create('-vol', vg, x)
    rt_id = apx.data_create('-sino', pg)
    cfg = astra.astra_dict(apx.fp_algo_name)
    cfg['VolumeDataId'], cfg['ProjectionDataId'], cfg['ProjectorId'] = v_id, rt_id, pid
    alg = astra.algorithm.create(cfg)
    astra.algorithm.run(alg)
    assert np.array_equal(apx.data_get(rt_id), p)
    astra.algorithm.delete(alg)
    apx.data_delete(rt_id)
    apx.data_delete(v_id)
    # CUDA tensors pass straight through
    pt = ap.gpu_fp(pg, vg, torch.from_numpy(x.astype(np.float32)).cuda(), pid)
    assert pt.is_cuda and np.array_equal(pt.cpu().numpy(), p)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_white_projection_class_vs_reference_module(golden, tag):
    from whitomo_b200.white_projection import WhiteProjection
    z = golden("white_small")
    n, na = int(z[tag + "_n"]), int(z[tag + "_n_angles"])
    wp = WhiteProjection(source=z[tag + "_source"], pixel_size=float(z[tag + "_pixel_size"]),
                         element_numbers=np.array([301, 302]), element_absorptions=z[tag + "_ea"],
                         energy_grid=z[tag + "_grid"], ph_size=(n, n), n_angles=na)      # sinogram=None works (B1)
    assert wp.proj_shape == (na, int(np.sqrt(2) * n + 1)) and wp.conc_shape == (2, n, n)
    wp.calc_sinogram_with_gt(z[tag + "_gt"])
    assert rel_l2(wp.sinogram, z[tag + "_sinogram"]) < TOL_OP
    c = z[tag + "_c"]
    f = wp.func(c)
    assert abs(f - float(z[tag + "_func"])) < 1e-4 * float(z[tag + "_func"])
    g = wp.grad(c)
    assert g.dtype == np.float64 and g.shape == c.shape and rel_l2(g, z[tag + "_grad"]) < TOL_OP
    I, exp_arg, exp, flat_fp = wp.FP_white(c)
    assert I.shape == (wp.proj_shape[0] * wp.proj_shape[1],) and flat_fp.shape == (2, I.shape[0], 1)
    assert rel_l2(I, z[tag + "_I"]) < TOL_OP and rel_l2(wp.calc_mu(exp), z[tag + "_mu"]) < TOL_OP
    conc, mu = wp.BP_white(wp.sinogram - I, exp)
    assert rel_l2(conc, z[tag + "_grad"]) < TOL_OP
    assert rel_l2(wp.FP(c[0]), oracle.fp(oracle.Geometry.standard(n, na), c[0])) < TOL_OP


def test_generic_barrier_method_on_device_tensors(golden, capsys):
    """the array-agnostic driver with CUDA-tensor closures over the engine reproduces the reference golden"""
    from whitomo_b200 import barrier_method as bm, white_rec
    from whitomo_b200.white_projection import WhiteProjection
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    wp = WhiteProjection(z["source"], float(z["pixel_size"]), np.array([301, 302]), z["ea"], z["grid"], (n, n), n_angles=na)
    wp.calc_sinogram_with_gt(z["gt"])
    ir = float(z["ir_beta"])
    dev = lambda a: torch.from_numpy(np.asarray(a, np.float32)).cuda()
    goal = (lambda x: wp.func_grad_dev(x)[0].item(), lambda x: wp.grad_dev(x))
    reg = (lambda x: ir * torch.abs(x[0] * x[1]), lambda x: ir * white_rec.calc_uneq_reg_grad(x))
    nonneg = (lambda x: -x, lambda x: -torch.ones_like(x), lambda x: torch.clamp(x, min=0))
    le_one = (lambda x: x - 1, lambda x: torch.ones_like(x), lambda x: torch.clamp(x, max=1))
    p = {k[2:]: z[k].item() for k in z.files if k.startswith("p_")}
    ans, _ = bm.barrier_method(dev(z["x0"]), goal, reg_dict={"r": reg}, ineq_dict={"a": nonneg, "b": le_one}, **p)
    assert rel_l2(ans.cpu().numpy(), z["ans"]) < TOL_REC


def test_white_barrier_stack_vs_reference_golden(golden):
    """fused device loop (K FP + K BP per iteration) == the reference's barrier_method run, slice by slice"""
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import solvers
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    grid = z["grid"]
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    sp = Spectrum(z["source"] / np.sum(z["source"] * grid[:, 1]), grid[:, 1], z["ea"], float(z["pixel_size"]))
    gt = torch.from_numpy(z["gt"].astype(np.float32)).cuda()
    meas, _, _ = P.white_fp(sp, gt, want_qmu=False, want_loss=False)
    p = {k[2:]: z[k].item() for k in z.files if k.startswith("p_")}
    # a 3-slice stack: slice 0 and 2 are the golden problem, slice 1 a different start
    x0 = np.stack([z["x0"], 0.5 * z["x0"] + 0.2, z["x0"]]).astype(np.float32)
    x, info = solvers.white_barrier_stack(P, sp, meas.expand(3, -1, -1).contiguous(), torch.from_numpy(x0).cuda(),
                                          ir_beta=float(z["ir_beta"]), **p)
    assert rel_l2(x[0].cpu().numpy(), z["ans"]) < TOL_REC
    assert torch.equal(x[0], x[2]) and not torch.equal(x[0], x[1])
    assert info["iters"] > 0 and torch.isfinite(info["loss"]).all()


def test_fused_loop_statistics_callback_has_the_reference_format(golden):
    """add_stat_cb of the fused loop receives (goal_stat, reg_stats, bf_stats) of StepStat(x, func, grad) like the
    reference's (src/barrier_method.py:182-199): goal func = f(x), goal grad = the TOTAL gradient (B5), the named
    regulariser and barrier entries of src/barrier_method_gogo.py:79-94 -- checked against the oracle's driver"""
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import solvers
    from whitomo_b200.barrier_method import StepStat
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    grid = z["grid"]
    wo = ow.WhiteOracle(z["source"], float(z["pixel_size"]), z["ea"], grid, (n, n), n_angles=na)
    wo.calc_sinogram_with_gt(z["gt"])
    ir_beta = float(z["ir_beta"])
    p = {k[2:]: z[k].item() for k in z.files if k.startswith("p_")}
    ref = []
    ineq_reg = (lambda x: ir_beta * np.abs(x[0] * x[1]), lambda x: ir_beta * ow.calc_uneq_reg_grad(x))
    nonneg = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    le_one = (lambda x: x - 1, lambda x: np.ones_like(x), lambda x: np.clip(x, x.min(), 1))
    ob.barrier_method(z["x0"].copy(), (wo.func, wo.grad), reg_dict={"ineq_reg": ineq_reg},
                      ineq_dict={"conc_non_negative": nonneg, "conc_less_than_one": le_one}, add_stat_cb=ref.append, **p)
    P = Projector(n, n, wo.geom.ndet, np.linspace(0, np.pi, na))
    sp = Spectrum(wo.source, grid[:, 1], z["ea"], float(z["pixel_size"]))
    meas = torch.from_numpy(wo.sinogram.reshape((1,) + wo.proj_shape).astype(np.float32)).cuda()
    seen = []
    solvers.white_barrier_stack(P, sp, meas, torch.from_numpy(z["x0"].astype(np.float32)).cuda()[None], ir_beta=ir_beta,
                                add_stat_cb=seen.append, **p)
    assert len(seen) == len(ref) > 3
    for i in (0, 1, len(ref) - 1):
        goal, regs, bfs = seen[i]
        f_ref, x_ref, g_ref = ref[i]
        assert isinstance(goal, StepStat) and set(regs) == {"ineq_reg"} and set(bfs) == {"conc_non_negative", "conc_less_than_one"}
        assert goal.x.slices.tolist() == [0]
        assert rel_l2(goal.x[0].cpu().numpy(), x_ref) < 1e-3 and abs(goal.func[0].item() - f_ref) < 1e-3 * abs(f_ref)
        assert rel_l2(goal.grad[0].cpu().numpy(), g_ref) < 2e-3                      # the total gradient
        xr = x_ref
        assert rel_l2(regs["ineq_reg"].func[0].cpu().numpy(), ir_beta * np.abs(xr[0] * xr[1])) < 2e-3
        assert rel_l2(bfs["conc_non_negative"].grad[0].cpu().numpy(), -1.0 / xr) < 2e-3
        assert rel_l2(bfs["conc_less_than_one"].func[0].cpu().numpy(), -np.log(1 - xr)) < 2e-3


def test_white_rec_stack_vs_oracle_iteration(golden):
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import solvers
    z = golden("white_small")
    n, na = int(z["a_n"]), int(z["a_n_angles"])
    grid = z["a_grid"]
    wo = ow.WhiteOracle(z["a_source"], float(z["a_pixel_size"]), z["a_ea"], grid, (n, n), n_angles=na)
    wo.calc_sinogram_with_gt(z["a_gt"])
    P = Projector(n, n, wo.geom.ndet, np.linspace(0, np.pi, na))
    sp = Spectrum(wo.source, grid[:, 1], z["a_ea"], float(z["a_pixel_size"]))
    meas = torch.from_numpy(wo.sinogram.reshape(wo.proj_shape).astype(np.float32)).cuda()
    c = z["a_c"].copy()
    alpha, beta = 0.7, 1e-2                                                   # src/white_rec.py:42-43
    ref_losses = []
    for it in range(9):
        c, fl = ow.iteration(wo, c, alpha, beta, clip_concentrations=True, dissimilarity_constraint=True,
                             triv_conc_constraint=True, iter_num=it)
        ref_losses.append(fl)
    cd, losses = solvers.white_rec_stack(P, sp, meas[None], torch.from_numpy(z["a_c"].astype(np.float32)).cuda(),
                                         n_iters=9, alpha=alpha, beta_reg=beta)
    assert rel_l2(cd[0].cpu().numpy(), c) < TOL_REC
    assert np.allclose([l.item() for l in losses], ref_losses, rtol=1e-3)


def test_mar_barrier_least_squares_vs_reference_golden(golden):
    from whitomo_b200.projector import Projector
    from whitomo_b200 import solvers
    z = golden("barrier_mar")
    n, na = int(z["n"]), int(z["n_angles"])
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    nb = int(z["n_biter"])
    x = solvers.barrier_least_squares(P, z["projections"].copy(), float(z["i0"]), float(z["bound"]), n_biter=nb)
    # measured 2.3e-6 / 2.1e-5 (profiles/r2_test_margins.json)
    assert x is not None and rel_l2(x.cpu().numpy(), z["ans"]) < TOL_REC
    x2 = solvers.barrier_least_squares(P, z["projections"].copy(), float(z["i0"]), float(z["bound"]), n_iter=100,
                                       n_biter=nb, t0=0.1, max_iter=1000, do_alpha_decay=True, with_sino_barrier=False)
    assert rel_l2(x2.cpu().numpy(), z["ans_no_ineq"]) < TOL_REC


def test_hough_barrier_class_on_device_closures(golden):
    from whitomo_b200.projector import Projector
    from whitomo_b200.barrier import HoughBarrier
    n, na = 32, 24
    g = oracle.Geometry.standard(n, na)
    P = Projector(n, n, g.ndet, np.linspace(0, np.pi, na))
    rng = np.random.default_rng(5)
    x = rng.uniform(0.2, 1.0, n * n).astype(np.float32)
    FPn = lambda v: oracle.fp(g, v.reshape(n, n)).flatten()
    m = (rng.random(na * g.ndet) < 0.1) & (FPn(x) > 5.0)          # constrained rays well inside the object
    BPn = lambda v: oracle.bp(g, v.reshape(g.sino_shape)).flatten()
    FPd = lambda v: P.fp(v.reshape(n, n)).flatten()
    BPd = lambda v: P.bp(v.reshape(g.sino_shape)).flatten()
    b = 0.5 * float(FPn(x)[m].min())
    hn = ob.HoughBarrier(FPn, BPn, m, b, np.zeros(g.sino_shape, np.float32))
    hd = HoughBarrier(FPd, BPd, torch.from_numpy(m).cuda(), b, torch.zeros(g.sino_shape, device="cuda"))
    pn, gn = hn.eval(x)
    pd, gd = hd.eval(torch.from_numpy(x).cuda())
    assert rel_l2(pd.cpu().numpy(), pn) < TOL_OP and rel_l2(gd.cpu().numpy(), gn) < TOL_OP
    assert hd.is_feasible(torch.from_numpy(x).cuda())
    x_bad = 0.01 * x
    assert not hd.is_feasible(torch.from_numpy(x_bad).cuda())
    assert rel_l2(hd.project(torch.from_numpy(x_bad).cuda()).cpu().numpy(), hn.project(x_bad)) < TOL_REC


def test_fp_angle_range_and_angle_split_single_rank():
    from whitomo_b200.projector import Projector
    from whitomo_b200.distributed import AngleSplit
    P = Projector(48, 48, 69, np.linspace(0, np.pi, 40))
    x = torch.rand(2, 48, 48, device="cuda")
    full = P.fp(x)
    part = P.fp(x, angle_range=(11, 29))
    assert torch.equal(part[:, 11:29], full[:, 11:29]) and part[:, :11].abs().sum() == 0 and part[:, 29:].abs().sum() == 0
    sp = AngleSplit(P)
    assert sp.range == (0, 40) and torch.equal(sp.fp(x), full)
    y = torch.rand(40, 69, device="cuda")
    assert torch.equal(sp.bp(y), P.bp(y))


def test_stack_reconstruction_in_resident_batches(golden):
    """C4-style job at toy size: 7 host-resident slices, batches of 3, two 'ranks' -- every slice must come out
    bit-identical to reconstructing the whole stack in one resident batch (slices are independent, and a slice's
    result does not depend on which float4 group it shares)."""
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import solvers
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    grid = z["grid"]
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    sp = Spectrum(z["source"] / np.sum(z["source"] * grid[:, 1]), grid[:, 1], z["ea"], float(z["pixel_size"]))
    rng = np.random.default_rng(3)
    S = 7
    gts = np.stack([np.clip(z["gt"] * rng.uniform(0.5, 1.0), 0, 1) for _ in range(S)]).astype(np.float32)
    meas = P.white_fp(sp, torch.from_numpy(gts).cuda(), want_qmu=False, want_loss=False)[0].cpu().pin_memory()
    x0 = torch.from_numpy(np.stack([z["x0"] * rng.uniform(0.8, 1.2) for _ in range(S)]).astype(np.float32)).pin_memory()
    p = dict(n_iter=4, n_biter=2, t0=0.1, t_step=0.5, alpha=0.3, ir_beta=0.05, max_iter=10, n_steps_without_progress=None)
    whole, _ = solvers.white_barrier_stack(P, sp, meas.cuda(), x0.cuda(), **p)
    parts = []
    for rank in range(2):
        out, (s0, s1) = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=3, rank=rank, world=2, **p)
        assert (s0, s1) == ((0, 4), (4, 7))[rank]
        parts.append(out)
    assert torch.equal(torch.cat(parts), whole.cpu())


def test_white_projection_real_data_branch():
    """angles + sinogram given (src/white_projection.py:35-36,48-58): D = sinogram width, volume side int((D-1)/sqrt 2),
    the caller's sinogram is normalised by the source integral (on a copy, B2)"""
    from whitomo_b200.white_projection import WhiteProjection
    ang = np.deg2rad(np.arange(0, 200, 0.5))[::8]                    # teeth_recon/teeth_io.py:52 style sampling
    D = 91
    rng = np.random.default_rng(2)
    sino = rng.uniform(0.2, 1.0, (ang.shape[0], D))
    keep = sino.copy()
    grid = np.array([[2.5, 10], [0.005, 0.005]]).T
    wp = WhiteProjection(source=np.array([2.0, 1.0]), pixel_size=1e-5, element_numbers=np.array([301, 302]),
                         element_absorptions=np.array([[4000, 302.0], [300, 2000]]), energy_grid=grid, ph_size=None,
                         sinogram=sino, angles=ang)
    w = int((D - 1) / np.sqrt(2))
    assert wp.ph_size == (w, w) and wp.proj_shape == (ang.shape[0], D) and wp.conc_shape == (2, w, w)
    assert np.array_equal(sino, keep)                                 # not modified in place
    assert np.allclose(wp.sinogram, keep.ravel() / np.sum(np.array([2.0, 1.0]) * grid[:, 1]))
    c = rng.uniform(0.05, 0.2, wp.conc_shape)
    g_geom = oracle.Geometry(w, w, D, ang)
    wo = ow.WhiteOracle(np.array([2.0, 1.0]), 1e-5, np.array([[4000, 302.0], [300, 2000]]), grid, (w, w), geometry=g_geom)
    wo.sinogram = wp.sinogram
    assert abs(wp.func(c) - wo.func(c)) < 1e-4 * wo.func(c)
    assert rel_l2(wp.grad(c), wo.grad(c)) < TOL_OP


def test_example_script_reference_style_equals_fused():
    """examples/barrier_gogo.py: the reference's call pattern (numpy closures + barrier_method) and the fused device
    loop reconstruct the same concentrations"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "barrier_gogo.py"), "64", "60"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "relative L2 between the two reconstructions" in r.stdout


def test_masked_least_squares_scipy_cg_over_the_drop_in_and_device_cgls():
    """teeth_recon/experiments.py:366-401: the reference's fmin_cg call runs unchanged over gpu_fp/gpu_bp and follows
    the same path as over the oracle; the device CGLS reaches the same minimum of the same functional."""
    import scipy.optimize as opt
    from whitomo_b200 import astra_proxy as ap, solvers
    n, na = 32, 40
    pg = ap.astra.create_proj_geom('parallel', 1.0, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    vg = ap.astra.create_vol_geom(n, n)
    pid = ap.astra.create_projector('linear', pg, vg)
    g = oracle.Geometry.standard(n, na)
    z_ph = np.zeros((n, n))
    yy, xx = np.mgrid[0:n, 0:n]
    z_ph[(xx - 15.5) ** 2 + (yy - 15.5) ** 2 <= 144] = 0.10358164
    z_ph[(xx - 18) ** 2 + (yy - 12) ** 2 <= 9] = 9.21754054
    rng = np.random.RandomState(12345)
    i0, bound = 1e9, 8.0
    pr = rng.poisson(lam=np.exp(np.log(i0) - oracle.fp(g, z_ph).astype(np.float64)).flatten()).astype('float32').reshape(g.sino_shape)
    r, m = ob.ray_transform_from_projection(pr.copy(), bound, i0)
    b, mf = r.flatten(), m.flatten()

    def closures(FP, BP):
        def cost(x):
            d = FP(x) - b
            d[mf] = 0.0
            return np.sum(d ** 2)

        def grad(x):
            d = FP(x) - b
            d[mf] = 0.0
            return 2.0 * BP(d)
        return cost, grad

    dev_c = closures(lambda x: ap.gpu_fp(pg, vg, x.reshape(n, n), pid).flatten(),
                     lambda x: ap.gpu_bp(pg, vg, x.reshape(g.sino_shape), pid).flatten())
    cpu_c = closures(lambda x: oracle.fp(g, x.reshape(n, n)).flatten(),
                     lambda x: oracle.bp(g, x.reshape(g.sino_shape)).flatten())
    x0 = np.zeros(n * n, dtype='float32')
    xd = opt.fmin_cg(dev_c[0], x0, dev_c[1], maxiter=8, disp=False)
    xc = opt.fmin_cg(cpu_c[0], x0, cpu_c[1], maxiter=8, disp=False)
    assert rel_l2(xd, xc) < 5e-3          # the line search amplifies the fp32-level operator differences
    assert abs(dev_c[0](xd) - cpu_c[0](xc)) < 1e-2 * cpu_c[0](xc)
    assert dev_c[0](xd) < 0.05 * dev_c[0](x0.astype(np.float64))
    # device CGLS: monotone cost, and after 200 iterations at least as low as 60 fmin_cg iterations on the CPU
    P = ap.astra.projector_for(pg, vg, pid)
    xs, costs = solvers.mask_linear_least_squares(P, pr.copy(), i0, bound, n_iters=200)
    assert all(c1 <= c0 * (1 + 1e-4) for c0, c1 in zip(costs, costs[1:]))
    x60 = opt.fmin_cg(cpu_c[0], x0, cpu_c[1], maxiter=60, disp=False)
    assert cpu_c[0](xs.cpu().numpy().flatten().astype(np.float64)) <= 1.05 * cpu_c[0](x60)


def test_active_set_compaction_is_bit_identical_and_saves_work(golden):
    """slices that leave a barrier level early are dropped from the working set until the next level: same result
    bit for bit as carrying them along masked, fewer slice-iterations spent"""
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import solvers
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    grid = z["grid"]
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    sp = Spectrum(z["source"] / np.sum(z["source"] * grid[:, 1]), grid[:, 1], z["ea"], float(z["pixel_size"]))
    rng = np.random.default_rng(11)
    S = 6
    gts = np.stack([np.clip(z["gt"] * s, 0, 1) for s in (1.0, 0.2, 0.9, 0.05, 0.6, 0.4)]).astype(np.float32)
    meas = P.white_fp(sp, torch.from_numpy(gts).cuda(), want_qmu=False, want_loss=False)[0]
    x0 = torch.from_numpy(np.stack([z["x0"] * rng.uniform(0.7, 1.3) for _ in range(S)]).astype(np.float32)).cuda()
    p = dict(n_iter=12, n_biter=3, t0=0.1, t_step=0.5, alpha=0.3, ir_beta=0.05, max_iter=30,
             n_steps_without_progress=2, min_delta_loss=0.5)
    xa, ia = solvers.white_barrier_stack(P, sp, meas, x0, compact=True, **p)
    xb, ib = solvers.white_barrier_stack(P, sp, meas, x0, compact=False, **p)
    assert torch.equal(xa, xb) and torch.equal(ia["loss"], ib["loss"])
    assert torch.equal(ia["slice_iters"], ib["slice_iters"])
    spent = int(ia["slice_iters"].sum().item())
    assert spent < ia["iters"] * S                      # some slices did stop early ...
    assert len(set(ia["slice_iters"].tolist())) > 1     # ... at different times


def test_soft_inequality_least_squares_scipy_cg_over_the_drop_in():
    """teeth_recon/experiments.py:513-582 (ineq_linear_least_squares): the reference's closures -- including its
    literal `d[1-m]` integer indexing, which only ever touches d[0] and d[1] -- handed to scipy.optimize.fmin_cg
    run unchanged over gpu_fp/gpu_bp and track the same closures over the oracle."""
    import scipy.optimize as opt
    from whitomo_b200 import astra_proxy as ap
    n, na = 32, 40
    pg = ap.astra.create_proj_geom('parallel', 1.0, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    vg = ap.astra.create_vol_geom(n, n)
    pid = ap.astra.create_projector('linear', pg, vg)
    g = oracle.Geometry.standard(n, na)
    ph = np.zeros((n, n))
    yy, xx = np.mgrid[0:n, 0:n]
    ph[(xx - 15.5) ** 2 + (yy - 15.5) ** 2 <= 144] = 0.10358164
    ph[(xx - 18) ** 2 + (yy - 12) ** 2 <= 9] = 9.21754054
    rng = np.random.RandomState(12345)
    i0, bound, alpha = 1e9, 8.0, 300.0
    pr = rng.poisson(lam=np.exp(np.log(i0) - oracle.fp(g, ph).astype(np.float64)).flatten()).astype('float32').reshape(g.sino_shape)
    r, m = ob.ray_transform_from_projection(pr.copy(), bound, i0)
    m = m.flatten()
    b = r.flatten()
    b[m] = np.log(i0) - np.log(bound)
    n_bad = np.sum(m)
    n_good = m.size - n_bad

    def closures(FP, BP):
        def weigh(x, good_w, bad_w):
            d = FP(x) - b
            d[(d > 0) * m] = 0.0
            d[m] *= alpha
            d[1 - m] /= good_w                 # the reference's literal indexing (elements 0 and 1 only)
            d[m] /= bad_w
            return d
        return (lambda x: np.sum(weigh(x, np.sqrt(n_good), np.sqrt(n_bad)) ** 2)), (lambda x: 2.0 * BP(weigh(x, n_good, n_bad)))

    dev_c = closures(lambda x: ap.gpu_fp(pg, vg, x.reshape(n, n), pid).flatten(),
                     lambda x: ap.gpu_bp(pg, vg, x.reshape(g.sino_shape), pid).flatten())
    cpu_c = closures(lambda x: oracle.fp(g, x.reshape(n, n)).flatten(),
                     lambda x: oracle.bp(g, x.reshape(g.sino_shape)).flatten())
    x0 = np.zeros(n * n, dtype='float32')
    xd = opt.fmin_cg(dev_c[0], x0, dev_c[1], maxiter=6, disp=False)
    xc = opt.fmin_cg(cpu_c[0], x0, cpu_c[1], maxiter=6, disp=False)
    assert rel_l2(xd, xc) < 5e-3
    assert dev_c[0](xd) < dev_c[0](x0.astype(np.float64))


def test_soft_inequality_least_squares_device_loop():
    """solvers.ineq_linear_least_squares (teeth_recon/experiments.py:513-582 as a device loop: one FP + one BP per
    iteration, no host synchronisation) on the B200 kernels against the same loop over the CPU oracle
    (tests/test_soft_ineq_cpu.py ties that loop to the reference's closures and to scipy's fmin_cg)"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_projector import OracleProjector
    from test_soft_ineq_cpu import _problem, reference_closures
    from whitomo_b200.projector import Projector
    from whitomo_b200 import solvers
    g, ph, pr, i0, bound, alpha = _problem()
    n = g.rows
    ang = np.linspace(0, np.pi, g.nangles)
    xc, hc = solvers.ineq_linear_least_squares(OracleProjector(n, n, g.ndet, ang), pr.copy(), i0, bound, alpha, n_iters=25)
    xd, hd = solvers.ineq_linear_least_squares(Projector(n, n, g.ndet, ang), pr.copy(), i0, bound, alpha, n_iters=25)
    assert np.allclose(hd.cpu().numpy()[:10], hc.numpy()[:10], rtol=1e-3)          # same descent, step by step
    assert rel_l2(xd.cpu().numpy(), xc.numpy()) < 2e-2                             # CG amplifies fp32 rounding slowly
    cost, grad = reference_closures(g, pr, i0, bound, alpha,
                                    lambda x: oracle.fp(g, x.reshape(n, n)).flatten().astype(np.float64),
                                    lambda x: oracle.bp(g, x.reshape(g.sino_shape)).flatten().astype(np.float64))
    x0 = np.zeros(n * n)
    assert np.linalg.norm(grad(xd.cpu().numpy().astype(np.float64).flatten())) < 0.05 * np.linalg.norm(grad(x0))
    xl, hl = solvers.ineq_linear_least_squares(Projector(n, n, g.ndet, ang), pr.copy(), i0, bound, alpha, n_iters=150)
    assert hl[-1].item() < 1e-4 * hl[0].item()


def test_stack_job_checkpoint_and_resume(golden, tmp_path, monkeypatch):
    """every finished batch is written to the checkpoint directory; a restarted job loads what it finds (bit-identical)
    and recomputes only what is missing or was written for other data"""
    import os
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import solvers
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    grid = z["grid"]
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    sp = Spectrum(z["source"] / np.sum(z["source"] * grid[:, 1]), grid[:, 1], z["ea"], float(z["pixel_size"]))
    rng = np.random.default_rng(5)
    S = 8
    gts = np.stack([np.clip(z["gt"] * rng.uniform(0.5, 1.0), 0, 1) for _ in range(S)]).astype(np.float32)
    meas = P.white_fp(sp, torch.from_numpy(gts).cuda(), want_qmu=False, want_loss=False)[0].cpu().pin_memory()
    x0 = torch.from_numpy(np.stack([z["x0"] * rng.uniform(0.8, 1.2) for _ in range(S)]).astype(np.float32)).pin_memory()
    p = dict(n_iter=3, n_biter=2, t0=0.1, t_step=0.5, alpha=0.3, ir_beta=0.05, max_iter=10, n_steps_without_progress=None)
    ck = str(tmp_path / "ck")
    ref, _ = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=3, **p)
    first, _ = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=3, checkpoint=ck, **p)
    assert torch.equal(first, ref)
    files = sorted(f for f in os.listdir(ck) if f.endswith(".npy"))
    assert len(files) == 3 and not [f for f in os.listdir(ck) if f.endswith(".tmp")]

    calls = []
    real = solvers.white_barrier_stack

    def counting(*a, **k):
        calls.append(a[2].shape[0])
        return real(*a, **k)
    monkeypatch.setattr(solvers, "white_barrier_stack", counting)
    again, _ = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=3, checkpoint=ck, **p)
    assert torch.equal(again, ref) and calls == []                       # everything came from the files
    os.remove(os.path.join(ck, files[1]))
    again, _ = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=3, checkpoint=ck, **p)
    assert torch.equal(again, ref) and calls == [3]                      # only the missing batch was recomputed
    # other parameters or other data never reuse these files
    calls.clear()
    p2 = dict(p, alpha=0.2)
    other, _ = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=3, checkpoint=ck, **p2)
    assert len(calls) == 3 and not torch.equal(other, ref)
    calls.clear()
    meas2 = meas.clone()
    meas2[0] *= 1.01
    solvers.white_barrier_reconstruct(P, sp, meas2, x0, batch=3, checkpoint=ck, **p)
    assert calls == [3]                                                   # the batch holding the changed slice
    # another assumed spectrum (same sizes) is another job; a statistics callback is not
    calls.clear()
    sp2 = Spectrum(z["source"] / np.sum(z["source"] * grid[:, 1]), grid[:, 1], 1.05 * z["ea"], float(z["pixel_size"]))
    solvers.white_barrier_reconstruct(P, sp2, meas, x0, batch=3, checkpoint=ck, **p)
    assert len(calls) == 3
    calls.clear()
    again, _ = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=3, checkpoint=ck, stat_cb=lambda st: None, **p)
    assert torch.equal(again, ref) and calls == [3]       # same job key: only the batch the meas2 run overwrote is redone
