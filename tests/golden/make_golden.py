#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ -- run in the BUILD container only
(it reads /root/reference, which does not exist on the GPU box; tests only read the .npz files).

Two kinds of fixtures:

1. ASTRA-made artifacts shipped by the reference, repackaged compactly (no arithmetic of ours):
     g1_teeth.npz     teeth_recon/phantom.txt (10x10) + teeth_recon/projjections.txt (15x35 = sino.T)
                      written by teeth_recon/experiments.py:227,244 (i0=1e9, 35 angles, D=15, Poisson)
     g2_button4.npz   testdata/whitereconstruct/correct_button4/c{1,2}.txt (bit-packed) +
                      src/fbp.nptxt (written by src/barrier_method_gogo_additional_plots.py:93-97)

2. Outputs of the reference's OWN Python modules, executed here unmodified except for a load-time
   py2->py3 patch (``.iteritems()`` -> ``.items()``), on an ``astra_proxy`` module whose
   gpu_fp/gpu_bp are backed by the CPU oracle (ASTRA itself is not installable):
     white_small.npz    src/white_projection.py  WhiteProjection.{calc_sinogram_with_gt, func, grad}
     barrier_white.npz  src/barrier_method.py    barrier_method on the white-beam goal with the gogo
                        box constraints / |c0 c1| regulariser (src/barrier_method_gogo.py:79-94,179-193)
     barrier_mar.npz    teeth_recon/experiments.py:651-749 barrier_least_squares (function source
                        extracted with ast and executed) + teeth_recon/barrier.py HoughBarrier

Usage:  python tests/golden/make_golden.py
"""
import ast
import io
import os
import sys
import types
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


# ----------------------------------------------------------------------------------------------
# 1. repackage ASTRA-made artifacts
# ----------------------------------------------------------------------------------------------
def repackage():
    ph = np.loadtxt(os.path.join(REF, "teeth_recon/phantom.txt"))
    pr = np.loadtxt(os.path.join(REF, "teeth_recon/projjections.txt"))
    np.savez_compressed(os.path.join(HERE, "g1_teeth.npz"), phantom=ph, projections_T=pr,
                        i0=1e9, n_angles=35, ndet=15)
    d = os.path.join(REF, "testdata/whitereconstruct/correct_button4")
    c1 = np.loadtxt(os.path.join(d, "c1.txt"))
    c2 = np.loadtxt(os.path.join(d, "c2.txt"))
    assert set(np.unique(c1)) <= {0.0, 1.0} and set(np.unique(c2)) <= {0.0, 1.0}
    fbp = np.loadtxt(os.path.join(REF, "src/fbp.nptxt"))
    assert np.array_equal(fbp, fbp.astype(np.float32).astype(np.float64))  # ASTRA output is float32
    np.savez_compressed(
        os.path.join(HERE, "g2_button4.npz"),
        c1_bits=np.packbits(c1.astype(np.uint8)), c2_bits=np.packbits(c2.astype(np.uint8)),
        shape=np.array(c1.shape), fbp=fbp.astype(np.float32),
        # src/phantoms.py:159-181 (get_synth_data_small)
        grid=np.array([[2.5, 10], [0.005, 0.005]]).T, source=np.array([2.0, 1.0]),
        element_absorptions=np.stack([np.array([4000, 302.04159662]), np.array([300, 2000])]),
        pixel_size=1e-6, n_angles=360)
    # 5-level 256^2 test image (external-code phantom; used as a parity input only)
    p = np.loadtxt(os.path.join(REF, "testdata/fht_sirt_results/phantom_line_256.txt"))
    side = int(round(np.sqrt(p.size)))
    np.savez_compressed(os.path.join(HERE, "phantom_line_256.npz"),
                        phantom=p.reshape(side, side).astype(np.float32))


# ----------------------------------------------------------------------------------------------
# 2. run the reference's own modules on an oracle-backed astra_proxy
# ----------------------------------------------------------------------------------------------
class _FakeAstra(object):
    """The three geometry helpers the reference calls (src/white_projection.py:41-63)."""

    @staticmethod
    def create_proj_geom(kind, det_width, det_count, angles):
        assert kind == "parallel"
        return {"type": kind, "DetectorWidth": det_width, "DetectorCount": det_count,
                "ProjectionAngles": np.asarray(angles)}

    @staticmethod
    def create_vol_geom(*a):
        if len(a) == 1:
            a = tuple(a[0])
        return {"GridRowCount": a[0], "GridColCount": a[1]}

    @staticmethod
    def create_projector(kind, pg, vg):
        assert kind == "linear"
        return 1


def _geom(pg, vg):
    return oracle.Geometry(vg["GridRowCount"], vg["GridColCount"], pg["DetectorCount"],
                           pg["ProjectionAngles"], pg["DetectorWidth"])


def install_fake_astra_proxy():
    m = types.ModuleType("astra_proxy")
    m.astra = _FakeAstra
    m.gpu_fp = lambda pg, vg, v, proj_id: oracle.fp(_geom(pg, vg), v)
    m.gpu_bp = lambda pg, vg, rt, proj_id, supersampling=1: oracle.bp(_geom(pg, vg), rt)
    m.__all__ = ["astra", "gpu_fp", "gpu_bp"]
    sys.modules["astra_proxy"] = m
    h = types.ModuleType("helpers")
    h.printProgressBar = lambda *a, **k: None
    sys.modules["helpers"] = h
    return m


def load_reference_module(name, relpath):
    src = open(os.path.join(REF, relpath)).read().replace(".iteritems()", ".items()")
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, relpath)
    sys.modules[name] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


class _SinoShim(object):
    """Lets the unmodified constructor run with 'no sinogram' (SURVEY B1): it calls
    sinogram.ravel() before its None check and derives ph_size from sinogram.shape[1]."""

    def __init__(self, n_angles, n):
        self.shape = (n_angles, n)

    def ravel(self):
        return np.zeros(1)


def discs(n, k=2):
    """Scaled two-disc phantom in the spirit of src/button_generator.py:24-41."""
    yy, xx = np.mgrid[0:n, 0:n]
    c = np.zeros((k, n, n))
    c[0][(xx - 0.3125 * n) ** 2 + (yy - 0.3125 * n) ** 2 <= (0.195 * n) ** 2] = 1
    c[1][(xx - 0.78125 * n) ** 2 + (yy - 0.78125 * n) ** 2 <= (0.078 * n) ** 2] = 1
    return c


def white_cases():
    wp_mod = load_reference_module("white_projection", "src/white_projection.py")
    out = {}
    rng = np.random.default_rng(12345)
    for tag, n, na, E in (("a", 32, 24, 2), ("b", 24, 17, 5)):
        if E == 2:
            grid = np.array([[2.5, 10], [0.005, 0.005]]).T
            source = np.array([2.0, 1.0])
            ea = np.stack([np.array([4000, 302.04159662]), np.array([300, 2000])])
        else:
            grid = np.stack([np.linspace(1, 9, E), rng.uniform(0.5, 1.5, E)]).T
            source = rng.uniform(0.2, 2.0, E)
            ea = rng.uniform(200, 4000, (2, E))
        pix = 1e-6 * 256 / n
        gt = discs(n)
        wp = wp_mod.WhiteProjection(source=source, pixel_size=pix, element_numbers=np.array([301, 302]),
                                    element_absorptions=ea, energy_grid=grid, ph_size=(n, n),
                                    sinogram=_SinoShim(na, n), n_angles=na)
        wp.calc_sinogram_with_gt(gt)
        c = rng.uniform(0.05, 0.25, gt.shape)
        f = wp.func(c)
        g = wp.grad(c)
        I, exp_arg, exp, flat_fp = wp.FP_white(c)
        mu = wp.calc_mu(exp)
        for k_, v_ in dict(n=n, n_angles=na, grid=grid, source=source, ea=ea, pixel_size=pix, gt=gt,
                           c=c, sinogram=wp.sinogram, func=f, grad=g, I=I, mu=mu).items():
            out["%s_%s" % (tag, k_)] = np.asarray(v_)
    np.savez_compressed(os.path.join(HERE, "white_small.npz"), **out)
    return wp_mod


def barrier_white_case(wp_mod):
    bm = load_reference_module("barrier_method", "src/barrier_method.py")
    # calc_uneq_reg_grad: src/white_rec.py:180-186 cannot be imported (import-time side effects,
    # SURVEY B14); extract just that function with ast and execute it.
    src = open(os.path.join(REF, "src/white_rec.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "calc_uneq_reg_grad"][0]
    ns = {"np": np, "uneq_reg_buffer": None}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "white_rec.py", "exec"), ns)
    calc_uneq = ns["calc_uneq_reg_grad"]

    n, na = 24, 20
    grid = np.array([[2.5, 10], [0.005, 0.005]]).T
    source = np.array([2.0, 1.0])
    ea = np.stack([np.array([4000, 302.04159662]), np.array([300, 2000])])
    pix = 1e-6 * 256 / n
    gt = discs(n)
    wp = wp_mod.WhiteProjection(source=source, pixel_size=pix, element_numbers=np.array([301, 302]),
                                element_absorptions=ea, energy_grid=grid, ph_size=(n, n),
                                sinogram=_SinoShim(na, n), n_angles=na)
    wp.calc_sinogram_with_gt(gt)
    goal = (lambda x: wp.func(x), lambda x: wp.grad(x))
    rng = np.random.default_rng(12345)
    x0 = rng.uniform(0.10, 0.20, gt.shape)
    # src/barrier_method_gogo.py:79-94
    ir_beta = 0.05
    ineq_reg = (lambda x: ir_beta * np.abs(x[0] * x[1]), lambda x: ir_beta * calc_uneq(x))
    nonneg = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    le_one = (lambda x: x - 1, lambda x: np.ones_like(x), lambda x: np.clip(x, x.min(), 1))
    # alpha=0.3 (gogo uses 1.0 at 256^2): at this tiny size alpha=1.0 overshoots onto the c=1 face in step 1
    params = dict(n_iter=8, n_biter=4, t0=0.1, t_step=0.5, beta_reg=1.0, alpha=0.3, max_iter=30,
                  n_steps_without_progress=15)
    with contextlib.redirect_stdout(io.StringIO()):
        ans, _ = bm.barrier_method(x0.copy(), goal, reg_dict={"ineq_reg": ineq_reg},
                                   ineq_dict={"conc_non_negative": nonneg, "conc_less_than_one": le_one},
                                   **params)
    np.savez_compressed(os.path.join(HERE, "barrier_white.npz"), n=n, n_angles=na, grid=grid,
                        source=source, ea=ea, pixel_size=pix, gt=gt, x0=x0, ans=ans, ir_beta=ir_beta,
                        **{"p_" + k: v for k, v in params.items()})
    return bm


MAR_N_BITER = 6


def barrier_mar_case(bm):
    # teeth_recon/barrier.py does sys.path.append + `import barrier_method` -> gets the patched module
    hb_mod = load_reference_module("barrier", "teeth_recon/barrier.py")
    src = open(os.path.join(REF, "teeth_recon/experiments.py")).read()
    want = ("barrier_least_squares", "ray_transform_from_projection", "vg_to_shape")
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want]
    # Load-time patch #2: the function hard-codes n_biter=10 / t_step=0.2; on any input we tried the
    # reference's own schedule shrinks t until a pixel is clipped to exactly 0 (outer iteration 8), the
    # next barrier gradient is -1/0 = -inf and the whole image turns NaN.  A NaN golden pins nothing,
    # so both runs are cut to n_biter=6 (every other statement of the function executes unchanged).
    seg = "\n\n".join(ast.get_source_segment(src, n) for n in fns)
    assert seg.count("n_biter=10") == 2
    fns = ast.parse(seg.replace("n_biter=10", "n_biter=%d" % MAR_N_BITER)).body
    fake = sys.modules["astra_proxy"]
    ns = {"np": np, "gpu_fp": fake.gpu_fp, "gpu_bp": fake.gpu_bp, "barrier_method": bm,
          "HoughBarrier": hb_mod.HoughBarrier, "stat_cb": None, "plot_stats": lambda *a: None,
          "opt_stats": []}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "experiments.py", "exec"), ns)

    n, na = 32, 40
    yy, xx = np.mgrid[0:n, 0:n]
    ph = np.zeros((n, n))
    c = (n - 1) / 2
    ph[(xx - c) ** 2 + (yy - c) ** 2 <= (0.375 * n) ** 2] = 0.10358164   # misc/calc_recon_areas.py:15
    ph[(xx - 0.56 * n) ** 2 + (yy - 0.375 * n) ** 2 <= (0.1 * n) ** 2] = 9.21754054
    pg = _FakeAstra.create_proj_geom("parallel", 1.0, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    vg = _FakeAstra.create_vol_geom(n, n)
    i0, bound = 1e9, 8.0
    rng = np.random.RandomState(12345)                                   # teeth_recon/experiments.py:269
    r = fake.gpu_fp(pg, vg, ph, 1)
    p = np.exp(np.log(i0) - r)                                           # :196-200
    p = rng.poisson(lam=p.flatten()).astype("float32").reshape(r.shape)
    with contextlib.redirect_stdout(io.StringIO()):
        ans, _, ans_no, _ = ns["barrier_least_squares"](i0, pg, vg, p.copy(), 1, bound, 0.1)
    np.savez_compressed(os.path.join(HERE, "barrier_mar.npz"), n=n, n_angles=na, phantom=ph,
                        projections=p, i0=i0, bound=bound, ans=ans, ans_no_ineq=ans_no,
                        n_biter=MAR_N_BITER)


if __name__ == "__main__":
    repackage()
    install_fake_astra_proxy()
    wp_mod = white_cases()
    bm = barrier_white_case(wp_mod)
    barrier_mar_case(bm)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
