"""The C4 job end to end on this rank's GPU: a host-resident stack of 2048^2 slices (1800 angles, K = 2, E = 64) goes
through solvers.white_barrier_reconstruct (resident batches, uploads overlapped with the iterations, results written
back to pinned host memory).  Reports slice-iterations/s of the whole call, wall-clocked around it -- counting the
iterations ACTUALLY spent per slice (slices that stop early spend fewer than the schedule's maximum) -- and the time the
full 2048-slice stack would take at that rate.  Under torchrun every rank takes its slice range (no collective).

  python tools/c4_job.py --slices 32 --batch 8 --n-iter 6 --n-biter 3
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from whitomo_b200 import phantoms, solvers, synthesis
from whitomo_b200.projector import Projector, Spectrum


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--angles", type=int, default=1800)
    ap.add_argument("--slices", type=int, default=32)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--n-iter", type=int, default=6)
    ap.add_argument("--n-biter", type=int, default=3)
    ap.add_argument("--alpha", type=float, default=0.003)
    ap.add_argument("--max-iter", type=int, default=0, help="cap of the per-level iteration count (default 4 x n-iter)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    n, na, S = args.n, args.angles, args.slices
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    grid, source, ea, pix = phantoms.synth_spectrum(64, n)
    sp = Spectrum(source / np.sum(source * grid[:, 1]), grid[:, 1], ea, pix)
    # measurements of the whole stack, synthesised on the device two slices at a time, parked in pinned host memory
    meas = torch.empty((S,) + P.sino_shape, dtype=torch.float32, pin_memory=True)
    x0 = torch.full((S, 2, n, n), 0.5, dtype=torch.float32).pin_memory()
    gts = torch.empty((S, 2, n, n), dtype=torch.float32)
    for s in range(0, S, 2):
        gt = torch.from_numpy(np.stack([phantoms.two_discs(n, z=(s + i) / S) for i in range(min(2, S - s))])).cuda()
        gts[s:s + gt.shape[0]] = gt.cpu()
        meas[s:s + gt.shape[0]] = synthesis.white_sinogram(P, sp, gt.reshape(-1, n, n)).reshape((-1,) + P.sino_shape).cpu()
    params = dict(n_iter=args.n_iter, n_biter=args.n_biter, alpha=args.alpha, max_iter=args.max_iter or 4 * args.n_iter)
    solvers.white_barrier_reconstruct(P, sp, meas[:4], x0[:4], batch=4, n_iter=2, n_biter=1, alpha=0.003)   # warm-up
    torch.cuda.synchronize()
    t = time.perf_counter()
    stats = {}
    out, (s0, s1) = solvers.white_barrier_reconstruct(P, sp, meas, x0, batch=args.batch, rank=rank, world=world, stats=stats,
                                                      **params)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    # iterations the schedule spends per slice when no slice stops early (n_iter doubles per level, t_step = 0.5)
    its, ni = 0, args.n_iter
    for _ in range(args.n_biter):
        its += ni
        ni = min(int(ni / 0.5), params["max_iter"])
    mine = gts[s0:s1]
    err0 = float(torch.linalg.vector_norm(x0[s0:s1] - mine) / torch.linalg.vector_norm(mine))
    err = float(torch.linalg.vector_norm(out - mine) / torch.linalg.vector_norm(mine))
    spent = stats["slice_iters"]
    done = int(spent.clamp(min=0).sum().item())          # slice-iterations really run (every one is K FP + K BP + K3)
    rate = done / dt
    print(json.dumps({"rank": rank, "world": world, "slices": [s0, s1], "batch": args.batch, "iterations_per_slice_max": its,
                      "iterations_per_slice_spent": {"min": int(spent.min().item()), "mean": float(spent.float().mean().item()),
                                                     "max": int(spent.max().item())},
                      "wall_s": dt, "slice_iterations_per_s": rate,
                      "gups": rate * 4 * n * n * na / 1e9,
                      "full_stack_2048_slices_same_mean_iterations_s": 2048 / world * float(spent.float().mean().item()) / rate,
                      "rel_l2_to_ground_truth_start": err0, "rel_l2_to_ground_truth_end": err,
                      "finite": bool(torch.isfinite(out).all()), "in_box": bool(((out >= 0) & (out <= 1)).all())}), flush=True)


if __name__ == "__main__":
    main()
