import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whitomo_b200 import phantoms, solvers, synthesis
from whitomo_b200.projector import Projector
n, na = int(sys.argv[1]), int(sys.argv[2])
P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
ph = torch.from_numpy(phantoms.tooth_phantom(n) * (256.0 / n)).cuda()
gen = torch.Generator(device="cuda").manual_seed(12345)
counts = synthesis.create_projection_with_poisson_noise(1e9, P, ph, generator=gen)
fbp = P.fbp(solvers.ray_transform_from_projection(counts.clone(), 8.0, 1e9)[0])
soft = ph < 0.5
print("fbp err soft %.4f" % (torch.linalg.vector_norm((fbp - ph)[soft]) / torch.linalg.vector_norm(ph[soft])).item())
for alpha in [float(a) for a in sys.argv[3:]]:
    k = [0]
    def cb(d):
        x = d["x"][0]
        print(n, "alpha", alpha, "outer", k[0], "t=%.1e" % d["t"], "finite", bool(torch.isfinite(x).all()), "min %.2e max %.2e" % (x.min().item(), x.max().item()),
              "cost %.3e" % d["stats"][0, 0].item(), "err soft %.4f all %.4f" % ((torch.linalg.vector_norm((x - ph)[soft]) / torch.linalg.vector_norm(ph[soft])).item(),
              (torch.linalg.vector_norm(x - ph) / torch.linalg.vector_norm(ph)).item()), flush=True)
        k[0] += 1
    t0 = time.time()
    solvers.barrier_least_squares(P, counts, 1e9, 8.0, n_iter=200, n_biter=6, alpha=alpha, on_outer=cb)
    torch.cuda.synchronize(); print("time %.1fs" % (time.time() - t0))
