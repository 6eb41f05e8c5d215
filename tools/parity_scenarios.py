#!/usr/bin/env python
"""Scenario-level parity at the BASELINE configurations, with the reference's own parameters: the device solver loops
against the oracle's restatement of the reference's loops, iterate by iterate.

  C2  src/white_rec.py:199-244,338-368   `Iteration` x 200 at 512^2 / 360 angles / E = 64 (alpha 0.7, beta 1e-2,
                                          clip + dissimilarity + trivial-concentration terms, element alternation)
  C3  teeth_recon/experiments.py:651-749 `barrier_least_squares`, first run, at 1024^2 / 720 angles with the reference's
                                          parameters (n_iter 200, n_biter 10, t0 0.2, t_step 0.2, alpha 0.1): every
                                          inner iteration until the reference's own arithmetic turns NaN
  C4  src/barrier_method_gogo.py:179-193  white-beam log-barrier run (alpha 1, t0 0.1, beta 1, 0.05 |c0 c1|) on one
                                          2048^2 / 1800-angle / E = 64 slice, first iterations

Writes the relative L2 distance of the iterates per iteration to profiles/r2_parity_scenarios.json.
Needs a B200; the oracle legs run on all host threads (C2 ~2 min, C3 ~1.5 min, C4 ~20 s per iteration).

    python tools/parity_scenarios.py [--scenarios C2,C3,C4] [--c4-iters 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _white_oracle(n, na, E, cls=None):
    import oracle
    from oracle import white as ow
    from whitomo_b200 import phantoms
    grid, source, ea, pix = phantoms.synth_spectrum(E, n)
    wo = (cls or ow.WhiteOracle)(source, pix, ea, grid, (n, n), n_angles=na)
    wo.FP = lambda x: oracle.fp(wo.geom, x, threads=0)      # all host threads: same arithmetic per ray
    wo.BP = lambda x: oracle.bp(wo.geom, x, threads=0)      # (per-thread partial volumes: summation order differs at 1e-7)
    return wo, grid, ea, pix


def scenario_c2(n_iters=200, checkpoints=(1, 2, 4, 8, 20, 40, 80, 120, 160, 200)):
    """white_rec `Iteration` x n_iters at 512^2 x 360 x E=64"""
    import torch
    from oracle import white as ow
    from whitomo_b200 import phantoms, solvers
    from whitomo_b200.projector import Projector, Spectrum
    n, na, E = 512, 360, 64
    wo, grid, ea, pix = _white_oracle(n, na, E)
    gt = phantoms.two_discs(n).astype(np.float64)
    wo.calc_sinogram_with_gt(gt)
    P = Projector(n, n, wo.geom.ndet, np.linspace(0, np.pi, na))
    sp = Spectrum(wo.source, grid[:, 1], ea, pix)
    # both sides reconstruct from the SAME measurement (the oracle's), as a user of the reference would
    meas = torch.from_numpy(wo.sinogram.reshape(wo.proj_shape).astype(np.float32)).cuda()
    rng = np.random.default_rng(12345)
    c0 = (0.1 + rng.uniform(-0.01, 0.01, gt.shape)).astype(np.float32)
    checkpoints = [k for k in checkpoints if k <= n_iters]
    t0 = time.perf_counter()
    c = c0.astype(np.float64)
    ref, ref_loss = {}, {}
    for it in range(n_iters):
        c, fl = ow.iteration(wo, c, 0.7, 1e-2, clip_concentrations=True, dissimilarity_constraint=True,
                             triv_conc_constraint=True, iter_num=it)
        if it + 1 in checkpoints:
            ref[it + 1], ref_loss[it + 1] = c.copy(), float(fl)
    t_cpu = time.perf_counter() - t0
    cd = torch.from_numpy(c0).cuda()
    out, done = [], 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in checkpoints:
        cd, losses = solvers.white_rec_stack(P, sp, meas, cd, n_iters=k - done, start_iter=done)
        done = k
        out.append({"iteration": k, "rel_l2": rel_l2(cd[0].cpu().numpy(), ref[k]),
                    "loss_gpu": float(losses[-1].item()), "loss_oracle_prev_iterate": ref_loss[k],
                    "rel_l2_to_ground_truth": rel_l2(cd[0].cpu().numpy(), gt)})
    t_gpu = time.perf_counter() - t0
    return {"config": "C2: 512^2 x 360 angles, K=2, E=64, white_rec Iteration, alpha 0.7, beta 1e-2", "n_iters": n_iters,
            "oracle_seconds": round(t_cpu, 1), "gpu_seconds_incl_checkpoint_downloads": round(t_gpu, 2), "iterates": out}


def scenario_c3(max_inner=9):
    """barrier_least_squares, first run, reference parameters, 1024^2 x 720"""
    import torch
    import oracle
    from oracle import barrier as ob
    from whitomo_b200 import phantoms, solvers
    from whitomo_b200.projector import Projector
    n, na = 1024, 720
    g = oracle.Geometry.standard(n, na)
    ph = phantoms.tooth_phantom(n) * (256.0 / n)        # pinned levels are mu*pixel at 256^2 sampling
    i0, bound = 1e9, 8.0
    r = oracle.fp(g, ph, threads=0)
    rng = np.random.RandomState(12345)                   # teeth_recon/experiments.py:269
    counts = rng.poisson(lam=np.exp(np.log(i0) - r).flatten()).astype("float32").reshape(r.shape)
    FP = lambda v: oracle.fp(g, v.reshape(n, n), threads=0).flatten()
    BP = lambda v: oracle.bp(g, v.reshape(g.sino_shape), threads=0).flatten()
    ref = []
    t0 = time.perf_counter()
    ob.barrier_least_squares(FP, BP, (n, n), counts.copy(), i0, bound, n_iter=max_inner, n_biter=1,
                             on_iter=lambda ib, ii, x: ref.append(x.reshape(n, n).copy()))
    t_cpu = time.perf_counter() - t0
    P = Projector(n, n, g.ndet, np.linspace(0, np.pi, na))
    out = []
    for k in range(1, max_inner + 1):
        x = solvers.barrier_least_squares(P, counts.copy(), i0, bound, n_iter=k, n_biter=1).cpu().numpy()
        xr = ref[k - 1]
        rec = {"inner_iteration": k,
               "oracle": {"min": float(np.nanmin(xr)) if np.isfinite(xr).any() else None, "zeros": int((xr == 0).sum()),
                          "inf": int(np.isinf(xr).sum()), "nan": int(np.isnan(xr).sum())},
               "gpu": {"zeros": int((x == 0).sum()), "inf": int(np.isinf(x).sum()), "nan": int(np.isnan(x).sum())}}
        fin = np.isfinite(xr) & np.isfinite(x)
        rec["rel_l2_on_finite_pixels"] = rel_l2(x[fin], xr[fin]) if fin.any() else None
        rec["same_nonfinite_pixels"] = bool(np.array_equal(np.isfinite(xr), np.isfinite(x)))
        out.append(rec)
    return {"config": "C3: 1024^2 x 720 angles, tooth phantom with a gold insert, i0 1e9, bound 8; barrier_least_squares "
                      "run 1 with the reference's parameters n_iter 200, n_biter 10, t0 0.2, t_step 0.2, alpha 0.1",
            "note": "the reference's own step size walks pixels onto x = 0 (clip) within a few iterations; the next "
                    "barrier gradient is -1/0 and the whole image turns NaN one iteration later -- on both sides",
            "oracle_seconds": round(t_cpu, 1), "iterates": out}


def scenario_c4(n_iters=10):
    """gogo white-beam barrier run on one C4 slice, first n_iters inner iterations"""
    import torch
    from oracle import barrier as ob
    from oracle import white as ow
    from whitomo_b200 import phantoms, solvers
    from whitomo_b200.projector import Projector, Spectrum
    n, na, E = 2048, 1800, 64
    wo, grid, ea, pix = _white_oracle(n, na, E, cls=ow.ChunkedWhiteOracle)
    gt = phantoms.two_discs(n).astype(np.float64)
    wo.calc_sinogram_with_gt(gt)
    rng = np.random.default_rng(12345)
    x0 = rng.uniform(0.10, 0.20, gt.shape)
    ir_beta = 0.05                                           # src/barrier_method_gogo.py:79-94
    ineq_reg = (lambda x: ir_beta * np.abs(x[0] * x[1]), lambda x: ir_beta * ow.calc_uneq_reg_grad(x))
    nonneg = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    le_one = (lambda x: x - 1, lambda x: np.ones_like(x), lambda x: np.clip(x, x.min(), 1))
    params = dict(n_iter=n_iters, n_biter=1, t0=0.1, t_step=0.5, beta_reg=1.0, alpha=1.0, max_iter=5000)
    ref = []
    t0 = time.perf_counter()
    # n_steps_without_progress = 15 cannot trigger within the first 15 iterations; None spares the oracle 2 of the
    # reference's 3 forward models per iteration without changing an iterate
    ob.barrier_method(x0.copy(), (wo.func, wo.grad), reg_dict={"ineq_reg": ineq_reg},
                      ineq_dict={"conc_non_negative": nonneg, "conc_less_than_one": le_one},
                      n_steps_without_progress=None, on_iter=lambda ib, ii, x: ref.append(x.copy()), **params)
    t_cpu = time.perf_counter() - t0
    P = Projector(n, n, wo.geom.ndet, np.linspace(0, np.pi, na))
    sp = Spectrum(wo.source, grid[:, 1], ea, pix)
    meas = torch.from_numpy(wo.sinogram.reshape((1,) + wo.proj_shape).astype(np.float32)).cuda()
    out = []
    for k in range(1, n_iters + 1):
        x, _ = solvers.white_barrier_stack(P, sp, meas, torch.from_numpy(x0.astype(np.float32)).cuda()[None],
                                           ir_beta=ir_beta, n_steps_without_progress=None,
                                           **dict(params, n_iter=k))
        out.append({"inner_iteration": k, "rel_l2": rel_l2(x[0].cpu().numpy(), ref[k - 1]),
                    "rel_l2_to_ground_truth": rel_l2(x[0].cpu().numpy(), gt)})
    return {"config": "C4: one 2048^2 slice x 1800 angles, K=2, E=64; barrier_method with the gogo parameters "
                      "(alpha 1, t0 0.1, beta_reg 1, 0.05 |c0 c1|, box barriers)", "oracle_seconds": round(t_cpu, 1),
            "iterates": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", default="C2,C3,C4")
    ap.add_argument("--c2-iters", type=int, default=200)
    ap.add_argument("--c4-iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_parity_scenarios.json"))
    args = ap.parse_args()
    res = {}
    for name in args.scenarios.split(","):
        if name == "C2":
            res[name] = scenario_c2(args.c2_iters)
        elif name == "C3":
            res[name] = scenario_c3()
        elif name == "C4":
            res[name] = scenario_c4(args.c4_iters)
        print(name, json.dumps(res[name]), flush=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
