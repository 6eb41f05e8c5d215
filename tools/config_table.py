"""GPU vs CPU-oracle timings at the reference-runnable configurations C1-C3 (the CPU legs are bounded samples of
the oracle port on the box's host cores; ASTRA's 2D CPU algorithms are single-threaded, so 1 thread is the faithful
number and all-threads the best-effort one).  Prints one JSON object; committed as profiles/r1_config_table.json."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import white as ow
from whitomo_b200 import phantoms, solvers, synthesis
from whitomo_b200.projector import Projector, Spectrum


def gpu_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def cpu_time(fn):
    t = time.perf_counter(); fn(); return time.perf_counter() - t


def main():
    nthreads = oracle.num_threads()
    out = {"host_threads": nthreads}
    for tag, n, na in (("C1", 256, 180), ("C2", 512, 360), ("C3", 1024, 720)):
        d = int(np.sqrt(2) * n + 1)
        g = oracle.Geometry(n, n, d, np.linspace(0, np.pi, na))
        P = Projector(n, n, d, np.linspace(0, np.pi, na))
        x = phantoms.two_discs(n)[0]
        xd = torch.from_numpy(x).cuda()
        pd_ = P.fp(xd)
        p = pd_.cpu().numpy()
        U = n * n * na
        row = {"N": n, "angles": na, "updates_per_op": U,
               "gpu_fp_ms": gpu_time(lambda: P.fp(xd)) * 1e3, "gpu_bp_ms": gpu_time(lambda: P.bp(pd_)) * 1e3,
               "cpu1_fp_s": cpu_time(lambda: oracle.fp(g, x)), "cpu1_bp_s": cpu_time(lambda: oracle.bp(g, p)),
               "cpuN_fp_s": cpu_time(lambda: oracle.fp(g, x, threads=0)), "cpuN_bp_s": cpu_time(lambda: oracle.bp(g, p, threads=0))}
        row["fp_speedup_vs_1_thread"] = row["cpu1_fp_s"] / (row["gpu_fp_ms"] * 1e-3)
        row["bp_speedup_vs_1_thread"] = row["cpu1_bp_s"] / (row["gpu_bp_ms"] * 1e-3)
        out[tag] = row
    # C1: SIRT iterations/s
    g = oracle.Geometry.standard(256, 180)
    P = Projector(256, 256, g.ndet, np.linspace(0, np.pi, 180))
    p = oracle.fp(g, phantoms.two_discs(256)[0])
    pd_ = torch.from_numpy(p).cuda()
    out["C1"]["gpu_sirt_its_per_s"] = 100 / gpu_time(lambda: P.sirt(pd_, 100), reps=3)
    out["C1"]["cpu1_sirt_its_per_s"] = 5 / cpu_time(lambda: oracle.sirt(g, p, 5))
    # C2: white_rec iterations/s (K = 2, E = 64)
    n, na = 512, 360
    grid, source, ea, pix = phantoms.synth_spectrum(64, n)
    wo = ow.WhiteOracle(source, pix, ea, grid, (n, n), n_angles=na)
    gt = phantoms.two_discs(n).astype(np.float64)
    wo.calc_sinogram_with_gt(gt)
    c = np.full(gt.shape, 0.1)
    out["C2"]["cpu1_white_rec_its_per_s"] = 1 / cpu_time(lambda: ow.iteration(wo, c, 0.7, 1e-2, True, True, True, 1))
    P = Projector(n, n, wo.geom.ndet, np.linspace(0, np.pi, na))
    sp = Spectrum(wo.source, grid[:, 1], ea, pix)
    meas = synthesis.white_sinogram(P, sp, torch.from_numpy(gt.astype(np.float32)).cuda())
    c0 = torch.full((2, n, n), 0.1, device="cuda")
    out["C2"]["gpu_white_rec_its_per_s"] = 200 / gpu_time(lambda: solvers.white_rec_stack(P, sp, meas, c0, n_iters=200), reps=2)
    # C3: MAR iterations/s (the reference spends >= 4 FP + 2 BP per iteration; timed here as such on the CPU)
    n, na = 1024, 720
    g = oracle.Geometry.standard(n, na)
    P = Projector(n, n, g.ndet, np.linspace(0, np.pi, na))
    ph = phantoms.tooth_phantom(n) * (256.0 / n)
    counts = synthesis.create_projection_with_poisson_noise(1e9, P, torch.from_numpy(ph).cuda(),
                                                            generator=torch.Generator(device="cuda").manual_seed(1))
    out["C3"]["gpu_mar_its_per_s"] = 400 / gpu_time(lambda: solvers.barrier_least_squares(P, counts, 1e9, 8.0, n_iter=200, n_biter=2, alpha=0.003), reps=1)
    t_fp = cpu_time(lambda: oracle.fp(g, ph, threads=0)); t_bp = cpu_time(lambda: oracle.bp(g, np.ones(g.sino_shape, np.float32), threads=0))
    out["C3"]["cpuN_mar_its_per_s_estimate"] = 1.0 / (4 * t_fp + 2 * t_bp)
    out["C3"]["cpu1_mar_its_per_s_estimate"] = 1.0 / (4 * out["C3"]["cpu1_fp_s"] + 2 * out["C3"]["cpu1_bp_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
