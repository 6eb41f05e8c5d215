#!/usr/bin/env python
"""A convergent reconstruction at C4 size: slices of the 2048^2 x 1800-angle white-beam problem (K = 2, E = 64) through
the fused log-barrier loop (solvers.white_barrier_stack, the gogo run of src/barrier_method_gogo.py:179-193) until the
error to the ground truth has fallen by more than 10x.  The reference's hard-coded alpha = 1 is tuned to its 256^2 / 360
angles run and is bang-bang at this size (profiles/r2_parity_scenarios.json, C4); the step size here is scaled by the
number of angles.  Prints / writes the error, data term and wall time per barrier level.

    python tools/c4_converge.py [--n 2048 --angles 1800 --slices 4 --alpha 0.2 --n-iter 50 --levels 8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from whitomo_b200 import phantoms, solvers
    from whitomo_b200.projector import Projector, Spectrum
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--angles", type=int, default=1800)
    ap.add_argument("--slices", type=int, default=4)
    ap.add_argument("--alpha", type=float, default=0.2)
    ap.add_argument("--n-iter", type=int, default=50)
    ap.add_argument("--max-iter", type=int, default=400)
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--t0", type=float, default=0.1)
    ap.add_argument("--no-decay", action="store_true", help="keep alpha constant over the levels (do_alpha_decay=False)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c4_converge.json"))
    args = ap.parse_args()
    n, na, S, K = args.n, args.angles, args.slices, 2
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    grid, source, ea, pix = phantoms.synth_spectrum(64, n)
    sp = Spectrum(source / np.sum(source * grid[:, 1]), grid[:, 1], ea, pix)
    gt = torch.from_numpy(np.stack([phantoms.two_discs(n, z=s / float(S)) for s in range(S)])).cuda()
    meas, _, _ = P.white_fp(sp, gt, want_qmu=False, want_loss=False)
    x = torch.full_like(gt, 0.5)
    rel = lambda v: float(torch.linalg.vector_norm((v - gt).double()) / torch.linalg.vector_norm(gt.double()))
    err0 = rel(x)
    _, _, l0 = P.white_fp(sp, x, meas=meas, want_intensity=False, want_qmu=False)
    rows = [{"level": -1, "iterations": 0, "rel_l2_to_ground_truth": err0, "data_term": float(l0.sum().item()), "wall_s": 0.0}]
    t, alpha, n_iter, total = args.t0, args.alpha, args.n_iter, 0
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    for lvl in range(args.levels):
        x, info = solvers.white_barrier_stack(P, sp, meas, x, n_iter=n_iter, n_biter=1, t0=t, t_step=0.5, alpha=alpha,
                                              beta_reg=1.0, ir_beta=0.05, n_steps_without_progress=None)
        total += int(info["slice_iters"].max().item())
        torch.cuda.synchronize()
        rows.append({"level": lvl, "t": t, "alpha": alpha, "iterations": total, "rel_l2_to_ground_truth": rel(x),
                     "data_term": float(info["loss"].sum().item()), "wall_s": time.perf_counter() - t_start})
        print(json.dumps(rows[-1]), flush=True)
        t *= 0.5                                           # src/barrier_method.py:231-236
        if not args.no_decay:
            alpha *= np.sqrt(0.5)
        n_iter = min(int(n_iter / 0.5), args.max_iter)
    res = {"config": "%d slices of %d^2 x %d angles, K=2, E=64, x0 = 0.5; white_barrier_stack, t0 %.3g, t_step 0.5, alpha0 %.3g, "
                     "0.05 |c0 c1|" % (S, n, na, args.t0, args.alpha),
           "error_reduction": err0 / rows[-1]["rel_l2_to_ground_truth"],
           "slice_iterations_per_s": S * total / rows[-1]["wall_s"], "levels": rows}
    print(json.dumps({k: v for k, v in res.items() if k != "levels"}))
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
