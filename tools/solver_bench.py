"""Iterations/s of the device solver loops at the reference-runnable configurations (development aid)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whitomo_b200 import phantoms, solvers, synthesis
from whitomo_b200.projector import Projector, Spectrum


def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t, r


def main():
    out = {}
    # C1: 256^2, 180 angles, 100 SIRT iterations
    n, na = 256, 180
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    p = P.fp(torch.from_numpy(phantoms.two_discs(n)[0]).cuda())
    P.sirt(p, 2)
    dt, _ = timed(lambda: P.sirt(p, 100))
    out["C1_sirt_100_its_ms"] = dt * 1e3
    # C2: 512^2, 360 angles, K=2, E=64, white_rec 200 iterations
    n, na = 512, 360
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    grid, source, ea, pix = phantoms.synth_spectrum(64, n)
    sp = Spectrum(source / np.sum(source * grid[:, 1]), grid[:, 1], ea, pix)
    gt = torch.from_numpy(phantoms.two_discs(n)).cuda()
    meas = synthesis.white_sinogram(P, sp, gt)
    c0 = torch.full_like(gt, 0.1)
    solvers.white_rec_stack(P, sp, meas, c0, n_iters=3)
    dt, _ = timed(lambda: solvers.white_rec_stack(P, sp, meas, c0, n_iters=200))
    out["C2_white_rec_200_its_ms"] = dt * 1e3
    out["C2_iters_per_s"] = 200 / dt
    # C3: 1024^2, 720 angles, MAR 2 x 200 iterations
    n, na = 1024, 720
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    ph = torch.from_numpy(phantoms.tooth_phantom(n) * (256.0 / n)).cuda()
    counts = synthesis.create_projection_with_poisson_noise(1e9, P, ph, generator=torch.Generator(device="cuda").manual_seed(1))
    solvers.barrier_least_squares(P, counts, 1e9, 8.0, n_iter=5, n_biter=1, alpha=0.003)          # warm-up
    dt, _ = timed(lambda: solvers.barrier_least_squares(P, counts, 1e9, 8.0, n_iter=200, n_biter=2, alpha=0.003))
    out["C3_mar_400_its_ms"] = dt * 1e3
    out["C3_iters_per_s"] = 400 / dt
    dt, _ = timed(lambda: solvers.barrier_least_squares(P, counts, 1e9, 8.0, n_iter=200, n_biter=2, alpha=0.003,
                                                        with_sino_barrier=False))
    out["C3_iters_per_s_without_sinogram_barrier"] = 400 / dt
    print(json.dumps(out))


if __name__ == "__main__":
    main()
