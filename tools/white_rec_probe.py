"""Where a C2 white_rec iteration (512^2 x 360, K = 2, E = 64) spends its time (development aid): host enqueue time of
solvers.white_rec_stack against the device time of the same call.
  python tools/white_rec_probe.py [iterations]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from whitomo_b200 import phantoms, solvers, synthesis
from whitomo_b200.projector import Projector, Spectrum

its = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n, na = 512, 360
grid, source, ea, pix = phantoms.synth_spectrum(64, n)
P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
sp = Spectrum(source / np.sum(source * grid[:, 1]), grid[:, 1], ea, pix)
gt = torch.from_numpy(phantoms.two_discs(n).astype(np.float32)).cuda()
meas = synthesis.white_sinogram(P, sp, gt)
c0 = torch.full((2, n, n), 0.1, device="cuda")
solvers.white_rec_stack(P, sp, meas, c0, n_iters=10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
solvers.white_rec_stack(P, sp, meas, c0, n_iters=its)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(json.dumps({"iterations": its, "host_enqueue_us_per_it": (t1 - t0) / its * 1e6,
                  "device_us_per_it": e0.elapsed_time(e1) / its * 1e3, "wall_us_per_it": (t2 - t0) / its * 1e6}))
