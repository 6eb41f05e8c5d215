#!/usr/bin/env python
"""The reference's headline experiment (src/barrier_method_gogo.py: white-beam reconstruction of two element maps
with log-barrier box constraints 0 <= c <= 1 and the 0.05*|c0 c1| regulariser) written against the drop-in modules.

Part 1 is the reference's own call pattern -- WhiteProjection + closures + barrier_method(...), numpy in/out -- and
runs unchanged on the sm_100a engine.  Part 2 is the fused device loop for the same problem.  Both print the
data-fidelity loss; the script exits non-zero if the two disagree.

    python examples/barrier_gogo.py [N] [n_angles]        (defaults: 128, 180; the reference uses 256, 360)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from whitomo_b200 import barrier_method, phantoms, solvers, white_rec        # noqa: E402
from whitomo_b200.white_projection import WhiteProjection                     # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    n_angles = int(sys.argv[2]) if len(sys.argv) > 2 else 180
    energy_grid, source, element_absorptions, pixel_size = phantoms.synth_spectrum(2, n)     # src/phantoms.py:159-181
    gt = phantoms.two_discs(n).astype(np.float64)                                            # button_4-like phantom

    wp = WhiteProjection(source=source, pixel_size=pixel_size, element_numbers=np.array([301, 302]),
                         element_absorptions=element_absorptions, ph_size=gt.shape[1:], n_angles=n_angles,
                         energy_grid=energy_grid)
    wp.calc_sinogram_with_gt(gt)
    goal = (lambda x: wp.func(x), lambda x: wp.grad(x))

    rng = np.random.default_rng(12345)
    concentrations = 0.15 + rng.uniform(-0.05, 0.05, gt.shape)                               # feasible start

    ir_beta = 0.05                                                                           # gogo :79-94
    ineq_reg = (lambda x: ir_beta * np.abs(x[0] * x[1]), lambda x: ir_beta * white_rec.calc_uneq_reg_grad(x))
    conc_non_negative = (lambda x: -x, lambda x: -np.ones_like(x), lambda x: np.clip(x, 0, x.max()))
    conc_less_than_one = (lambda x: x - 1, lambda x: np.ones_like(x), lambda x: np.clip(x, x.min(), 1))
    params = dict(n_iter=10, n_biter=4, t0=0.1, t_step=0.5, beta_reg=1.0, alpha=1.0, max_iter=40,
                  n_steps_without_progress=15)

    print("loss at start: %.6g" % goal[0](concentrations))
    t = time.perf_counter()
    stdout, sys.stdout = sys.stdout, open(os.devnull, "w")        # the driver prints a line per barrier level
    try:
        ans, _ = barrier_method.barrier_method(concentrations.copy(), goal, reg_dict={'ineq_reg': ineq_reg},
                                               ineq_dict={'conc_non_negative': conc_non_negative,
                                                          'conc_less_than_one': conc_less_than_one}, **params)
    finally:
        sys.stdout.close()
        sys.stdout = stdout
    t_generic = time.perf_counter() - t
    print("reference-style driver : loss %.6g   (%.2f s, numpy closures, 3K FP + K BP per iteration)"
          % (goal[0](ans), t_generic))

    proj, spec = wp._proj, wp._spectrum
    meas = wp._meas()[None]
    x0 = torch.from_numpy(concentrations.astype(np.float32)).cuda()[None]
    torch.cuda.synchronize()
    t = time.perf_counter()
    x, info = solvers.white_barrier_stack(proj, spec, meas, x0, ir_beta=ir_beta, **params)
    torch.cuda.synchronize()
    t_fused = time.perf_counter() - t
    print("fused device loop      : loss %.6g   (%.3f s, %d iterations, K FP + K BP per iteration)"
          % (info["loss"][0].item(), t_fused, info["iters"]))
    err = np.linalg.norm(x[0].cpu().numpy() - ans) / np.linalg.norm(ans)
    print("relative L2 between the two reconstructions: %.2e" % err)
    return 0 if err < 1e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
