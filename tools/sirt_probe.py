"""Where a small SIRT iteration spends its time (development aid): host enqueue time of the C loop against the
device time of the same call, at C1 (256^2 x 180) or any N,A.
  python tools/sirt_probe.py [N A [iterations]]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from whitomo_b200.projector import Projector

n, a = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 180)
its = int(sys.argv[3]) if len(sys.argv) > 3 else 500
d = int(np.sqrt(2) * n + 1)
P = Projector(n, n, d, np.linspace(0, np.pi, a))
x = torch.rand(1, n, n, device="cuda")
y = P.fp(x)
P.sirt(y, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
P.sirt(y, its)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(json.dumps({"N": n, "A": a, "iterations": its, "host_enqueue_us_per_it": (t1 - t0) / its * 1e6,
                  "device_us_per_it": e0.elapsed_time(e1) / its * 1e3, "wall_us_per_it": (t2 - t0) / its * 1e6}))
