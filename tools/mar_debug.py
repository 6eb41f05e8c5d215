import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whitomo_b200 import phantoms, solvers, synthesis, _capi
from whitomo_b200.projector import Projector
n, na = int(sys.argv[1]), int(sys.argv[2])
P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
ph = torch.from_numpy(phantoms.tooth_phantom(n) * (256.0 / n)).cuda()
gen = torch.Generator(device="cuda").manual_seed(12345)
i0, bound = 1e9, 8.0
p = synthesis.create_projection_with_poisson_noise(i0, P, ph, generator=gen)[None]
def fin(name, t): print(name, "finite", bool(torch.isfinite(t).all()), "min %.4g max %.4g" % (t.min().item(), t.max().item()), flush=True)
fin("counts", p)
r, m = solvers.ray_transform_from_projection(p, bound, i0)
bl = math.log(i0) - math.log(bound)
b = torch.where(m, torch.full_like(r, bl), r).contiguous(); b_hb = 0.95 * bl
inv = 1.0 / (m[0].numel() - m.reshape(1, -1).sum(1)).to(torch.float32)
m8 = m.to(torch.uint8).contiguous()
fin("b", b); print("inv_ngood", inv, "nmetal", int(m.sum()))
x = torch.full((1, n, n), 1e-2, device="cuda")
u, pr, st = P.mar_fp(x, b, m8, inv, 0.2, b_hb, want_proj=True)
fin("u0", u); fin("proj0", pr); print("st0", st)
x = solvers._bc_project(P, x, m, b_hb); x.clamp_(min=0)
fin("x after project", x)
u, pr, st = P.mar_fp(x, b, m8, inv, 0.2, b_hb, want_proj=True)
fin("u1", u); fin("proj1", pr); print("st1", st)
for it in range(6):
    g = P.bp(u); fin("g%d" % it, g)
    P.barrier_update(x, g, 0.1, 1.0, 0.2, lower_only=True, mode=_capi.UPD_STEP_ONLY); fin("x%d" % it, x)
    u, pr, st = P.mar_fp(x, b, m8, inv, 0.2, b_hb, want_proj=True); fin("u", u); print("st", st)
    if (x < 0).any():
        P.barrier_update(x, None, 0.0, 1.0, 0.0, lower_only=True, mode=_capi.UPD_CLIP_ONLY); fin("x clipped", x)
        u, pr, st = P.mar_fp(x, b, m8, inv, 0.2, b_hb, want_proj=True)
