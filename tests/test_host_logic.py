"""CPU-only checks of the product's host side: the C-ABI library loads and exports every declared symbol,
the `astra` shim builds the reference's geometry dicts, and the py3 host mirrors of barrier_method /
HoughBarrier / white_rec reproduce the golden outputs of the reference's own modules when their FP/BP
closures are backed by the oracle (the product never imports the oracle; these tests inject it)."""
import ctypes
import os
import re
import pickle

import numpy as np
import pytest

import oracle
from oracle import white as ow, barrier as ob
from conftest import rel_l2, ROOT


def test_capi_loads_and_exports_every_declared_symbol():
    from whitomo_b200 import _capi
    lib = _capi.load()
    hdr = open(os.path.join(ROOT, "include", "whitomo_b200.h")).read()
    declared = set(re.findall(r"\b(wt_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.wt_version() >= 100
    assert ctypes.sizeof(_capi.UpdateParams) == 32
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.wt_geom_dims(None, None, None, None, None) == -1
    assert b"null geometry" in lib.wt_last_error()
    assert lib.wt_workspace_bytes(None, 4) == 0


def test_no_cpu_fallback_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from whitomo_b200._capi import EngineError
    from whitomo_b200.projector import Projector
    with pytest.raises(EngineError):
        Projector(8, 8, 12, np.linspace(0, np.pi, 4))
    from whitomo_b200 import astra_proxy as ap
    pg = ap.astra.create_proj_geom("parallel", 1.0, 12, np.linspace(0, np.pi, 4))
    vg = ap.astra.create_vol_geom(8, 8)
    with pytest.raises(EngineError):
        ap.gpu_fp(pg, vg, np.zeros((8, 8)), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "whitomo_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f


def test_astra_shim_geometry_dicts():
    from whitomo_b200 import astra_proxy as ap
    astra = ap.astra
    ang = np.linspace(0, np.pi, 7)
    pg = astra.create_proj_geom("parallel", 1.0, 15, ang)
    assert pg["type"] == "parallel" and pg["DetectorCount"] == 15 and pg["DetectorWidth"] == 1.0
    assert np.array_equal(pg["ProjectionAngles"], ang)
    for vg in (astra.create_vol_geom(10, 12), astra.create_vol_geom((10, 12)), astra.create_vol_geom(*(10, 12))):
        assert vg["GridRowCount"] == 10 and vg["GridColCount"] == 12
        assert vg["option"]["WindowMinX"] == -6 and vg["option"]["WindowMaxY"] == 5
    assert astra.create_vol_geom(9)["GridColCount"] == 9
    pickle.loads(pickle.dumps((pg, vg)))                      # teeth_recon/experiments.py:284-285
    with pytest.raises(NotImplementedError):
        astra.create_proj_geom("fanflat", 1.0, 15, ang, 1, 1)
    p2 = ap.AstraProxy(2)
    assert (p2.fp_algo_name, p2.bp_algo_name, p2.sirt_algo_name, p2.fbp_algo_name) == ("FP", "BP", "SIRT", "FBP")
    assert p2.dart_mask_name == "DARTMASK"
    with pytest.raises(NotImplementedError):
        ap.AstraProxy(4)
    assert astra.astra_dict("FP") == {"type": "FP"}
    i = astra.data2d.create("-vol", vg, np.ones((10, 12)))
    assert astra.data2d.get(i).dtype == np.float32
    with pytest.raises(ValueError):
        astra.data2d.create("-sino", pg, np.ones((3, 3)))
    astra.data2d.delete(i)


def test_barrier_method_mirror_vs_reference_golden_white(golden, capsys):
    from whitomo_b200 import barrier_method as bm, white_rec
    z = golden("barrier_white")
    n, na = int(z["n"]), int(z["n_angles"])
    wo = ow.WhiteOracle(z["source"], float(z["pixel_size"]), z["ea"], z["grid"], (n, n), n_angles=na)
    wo.calc_sinogram_with_gt(z["gt"])
    ir = float(z["ir_beta"])
    reg = (lambda x: ir * np.abs(x[0] * x[1]), lambda x: ir * white_rec.calc_uneq_reg_grad(x))
    nonneg = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    le_one = (lambda x: x - 1, lambda x: np.ones_like(x), lambda x: np.clip(x, x.min(), 1))
    p = {k[2:]: z[k].item() for k in z.files if k.startswith("p_")}
    seen = []
    ans, stats = bm.barrier_method(z["x0"].copy(), (wo.func, wo.grad), reg_dict={"ineq_reg": reg},
                                   ineq_dict={"conc_non_negative": nonneg, "conc_less_than_one": le_one},
                                   add_stat_cb=seen.append, **p)
    assert stats == []
    assert rel_l2(ans, z["ans"]) < 1e-12
    goal, regs, bfs = seen[0]
    assert set(regs) == {"ineq_reg"} and set(bfs) == {"conc_non_negative", "conc_less_than_one"}
    assert goal.x.shape == z["x0"].shape and np.isscalar(float(goal.func))


def test_barrier_method_infeasible_start_returns_none(capsys):
    from whitomo_b200 import barrier_method as bm
    never = (lambda x: x + 1.0, lambda x: np.ones_like(x), lambda x: x)        # x + 1 <= 0 is never met by proj
    x, stats = bm.barrier_method(np.ones(4), (lambda x: 0.0, lambda x: np.zeros_like(x)), ineq_dict={"c": never},
                                 n_iter=1, n_biter=1)
    assert stats is None and np.array_equal(x, np.ones(4))
    assert "feasibility projection unsuccesfull" in capsys.readouterr().out


def test_hough_barrier_mirror_vs_reference_golden_mar(golden, capsys):
    from whitomo_b200 import barrier_method as bm
    from whitomo_b200.barrier import HoughBarrier
    z = golden("barrier_mar")
    n, na = int(z["n"]), int(z["n_angles"])
    g = oracle.Geometry.standard(n, na)
    FP = lambda x: oracle.fp(g, x.reshape(n, n)).flatten()
    BP = lambda x: oracle.bp(g, x.reshape(g.sino_shape)).flatten()
    i0, bound = float(z["i0"]), float(z["bound"])
    r, m = ob.ray_transform_from_projection(z["projections"].copy(), bound, i0)
    m, b = m.flatten(), r.flatten()
    bound_log = np.log(i0) - np.log(bound)
    b[m] = bound_log
    bound_log *= 0.95
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

    hb = HoughBarrier(FP, BP, m, bound_log, r)
    morezero = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    x0 = 1e-2 + np.zeros((n, n), dtype="float32").flatten()
    ans, _ = bm.barrier_method(x0, (cost, grad), ineq_dict={"bound": hb, "morezero": morezero}, n_iter=200,
                               n_biter=int(z["n_biter"]), t0=0.2, t_step=0.2, beta_reg=1.0, alpha=0.1, max_iter=400,
                               do_alpha_decay=False)
    assert rel_l2(ans.reshape(n, n), z["ans"]) < 1e-6
    assert hb.num_feval > 0 and hb.num_geval > 0


def test_white_rec_helpers_match_oracle():
    from whitomo_b200 import white_rec
    c = np.random.default_rng(0).uniform(-0.2, 1.2, (2, 9, 11))
    assert np.array_equal(white_rec.calc_uneq_reg_grad(c), ow.calc_uneq_reg_grad(c))
    assert np.allclose(white_rec.calc_triv_conc_grad(c), ow.calc_triv_conc_grad(c), rtol=1e-15)
    torch = pytest.importorskip("torch")
    ct = torch.from_numpy(c)
    assert np.array_equal(white_rec.calc_uneq_reg_grad(ct).numpy(), ow.calc_uneq_reg_grad(c))


def test_synthetic_inputs():
    from whitomo_b200 import phantoms
    grid, src, ea, pix = phantoms.synth_spectrum(2, 256)
    assert np.array_equal(grid, np.array([[2.5, 0.005], [10, 0.005]])) and pix == 1e-6     # src/phantoms.py:159-181
    grid, src, ea, pix = phantoms.synth_spectrum(64, 512)
    assert grid.shape == (64, 2) and np.isclose(grid[8, 0], 2.5) and np.isclose(grid[38, 0], 10.0)
    assert np.allclose(ea[:, 8], [4000, 300]) and np.allclose(ea[:, 38], [302.04159662, 2000])
    c = phantoms.two_discs(256)
    assert c.shape == (2, 256, 256) and abs(c[0].sum() - 7845) < 60 and abs(c[1].sum() - 1257) < 30
    t = phantoms.tooth_phantom(64)
    assert np.allclose(np.unique(t), [0, phantoms.CA_LEVEL, phantoms.AU_LEVEL])


def _plan(rows, cols, ndet, angles, det_width=1.0):
    from whitomo_b200 import _capi
    lib = _capi.load()
    ang = np.ascontiguousarray(np.asarray(angles, dtype=np.float64).astype(np.float32))
    cap = len(ang)
    ip = ctypes.POINTER(ctypes.c_int)
    first, count, vert, order = (np.zeros(cap, np.int32) for _ in range(4))
    nt, nph = ctypes.c_int(), ctypes.c_int()
    ps = np.zeros(5, np.int32)
    rc = lib.wt_geom_plan(rows, cols, ndet, det_width, ang.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(ang), cap,
                          first.ctypes.data_as(ip), count.ctypes.data_as(ip), vert.ctypes.data_as(ip),
                          order.ctypes.data_as(ip), ctypes.byref(nt), ctypes.byref(nph), ps.ctypes.data_as(ip))
    assert rc == 0, lib.wt_last_error()
    n = nt.value
    return first[:n], count[:n], vert[:n], order[:n], nph.value, ps


@pytest.mark.parametrize("n,na", [(10, 35), (256, 180), (1024, 720), (2048, 1800), (8192, 4096), (4096, 300)])
def test_forward_projector_plan_invariants(n, na):
    """host-side planning of K1 (no CUDA involved): the angle tiles partition the angle list in order, hold at most 8
    angles of one orientation whose rays stay within the staged window, and the launch order is a permutation of the
    tiles grouped by orientation"""
    ang = np.linspace(0, np.pi, na)
    d = int(np.sqrt(2) * n + 1)
    first, count, vert, order, nph, ps = _plan(n, n, d, ang)
    assert first[0] == 0 and np.array_equal(first[1:], (first + count)[:-1]) and first[-1] + count[-1] == na
    assert count.min() >= 1 and count.max() <= 8
    a32 = ang.astype(np.float32).astype(np.float64)
    vertical = np.abs(np.float32(np.sin(a32))) < np.abs(np.float32(np.cos(a32)))          # |ray_x| < |ray_y| on float32 values
    for f, c, v in zip(first, count, vert):
        assert (vertical[f:f + c] == bool(v)).all()
    # adequately sampled geometries (A >= N) keep their tiles (almost) full; under-sampled ones fall back to small tiles
    if na >= n and na >= 180:
        assert count.mean() >= 6.0
    if na * 8 < n:
        assert count.mean() < 6.0
    # launch order: permutation, one phase per orientation, phases contiguous
    assert sorted(order.tolist()) == list(range(len(first)))
    assert nph in (1, 2) and ps[0] == 0 and ps[nph] == len(first)
    for ph in range(nph):
        tiles = order[ps[ph]:ps[ph + 1]]
        assert len(set(vert[tiles].tolist())) == 1
        assert np.array_equal(tiles, np.sort(tiles))             # angle order inside a phase


def test_plan_rejects_bad_geometries_without_cuda():
    from whitomo_b200 import _capi
    lib = _capi.load()
    ang = np.zeros(4, np.float32)
    fp = ang.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    nt = ctypes.c_int()
    none = ctypes.POINTER(ctypes.c_int)()
    assert lib.wt_geom_plan(0, 8, 8, 1.0, fp, 4, 0, none, none, none, none, ctypes.byref(nt), none, none) == -1
    assert lib.wt_geom_plan(8, 8, 8, 0.5, fp, 4, 0, none, none, none, none, ctypes.byref(nt), none, none) == -2
    assert b"det_width" in lib.wt_last_error()
    assert lib.wt_geom_plan(8, 8, 8, 1.0, fp, 4, 0, none, none, none, none, ctypes.byref(nt), none, none) == 0 and nt.value >= 1


def test_data_calls_reject_bad_arguments_before_touching_cuda():
    """argument checks of the data entry points come first: no GPU needed to see WT_ERR_INVALID (-1)"""
    from whitomo_b200 import _capi
    lib = _capi.load()
    assert lib.wt_pair_norm2(None, 1, 16, None, None) == -1 and b"wt_pair_norm2" in lib.wt_last_error()
    buf = (ctypes.c_double * 32)()
    x = (ctypes.c_float * 32)()
    assert lib.wt_pair_norm2(ctypes.addressof(x), 0, 16, ctypes.addressof(buf), None) == -1
    assert lib.wt_pair_norm2(ctypes.addressof(x), 1, 0, ctypes.addressof(buf), None) == -1
    assert lib.wt_fp(None, None, None, 1, 0, 1, None, 0, None) == -1 and b"wt_fp" in lib.wt_last_error()
    assert lib.wt_barrier_update(None, None, 1, 1, 16, None, None, None, None, 0, None) == -1


@pytest.mark.parametrize("n,na,ndblk", [(64, 48, 3), (256, 180, 12), (2048, 1800, 91)])
def test_launch_order_visits_every_cta_once(n, na, ndblk):
    """fp.cuh: fp_block maps the linear block index to (tile, detector block): phase by phase, detector blocks slowest
    inside a phase.  Restated here on the planned order: the map must be a bijection onto tiles x detector blocks and
    keep all tiles of a detector block adjacent (that adjacency is what keeps the lines in flight L2-resident)."""
    first, count, vert, order, nph, ps = _plan(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    ntiles = len(first)

    def fp_block(lin0):
        s0, s1 = 0, ps[1]
        for ph in range(1, 4):
            if ph < nph and lin0 >= ps[ph] * ndblk:
                s0, s1 = ps[ph], ps[ph + 1]
        lin, nt = lin0 - s0 * ndblk, s1 - s0
        dblk = lin // nt
        return int(order[s0 + lin - dblk * nt]), int(dblk)

    seen = [fp_block(i) for i in range(ntiles * ndblk)]
    assert len(set(seen)) == ntiles * ndblk
    assert all(0 <= t < ntiles and 0 <= b < ndblk for t, b in seen)
    # consecutive CTAs share the detector block except at block / phase boundaries
    changes = sum(1 for (t0, b0), (t1, b1) in zip(seen, seen[1:]) if b0 != b1)
    assert changes == nph * ndblk - 1


@pytest.mark.parametrize("n,na,fp_expect,bp_expect,th_expect", [(256, 180, 2, 4, 16), (128, 90, 2, 2, 16), (384, 270, 2, 4, 16),
                                                              (512, 360, 2, 2, 16), (768, 540, 1, 1, 32), (1024, 720, 1, 1, 32), (2048, 1800, 1, 1, 32)])
def test_small_problems_are_split_over_clusters(n, na, fp_expect, bp_expect, th_expect):
    """host-side choice of the cluster split (no CUDA involved): only geometries that would leave SMs idle are split,
    every rank keeps at least two pipeline stages, and volumes from 768^2 up run unsplit"""
    from whitomo_b200 import _capi
    lib = _capi.load()
    d = int(np.sqrt(2) * n + 1)
    first = _plan(n, n, d, np.linspace(0, np.pi, na))[0]
    fs, bs, th = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.wt_split_plan(n, n, d, na, len(first), ctypes.byref(fs), ctypes.byref(bs), ctypes.byref(th)) == 0
    assert (fs.value, bs.value, th.value) == (fp_expect, bp_expect, th_expect)
    assert lib.wt_split_plan(0, n, d, na, 1, None, None, None) == -1
