"""Quick FP/BP throughput table on one GPU (development aid; bench.py is the contract)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from whitomo_b200.projector import Projector


def timeit(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    cfgs = [(256, 180, 8), (512, 360, 8), (1024, 720, 4), (2048, 1800, 1), (2048, 1800, 4), (2048, 1800, 8)]
    if len(sys.argv) > 1:
        cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for n, a, imgs in cfgs:
        d = int(np.sqrt(2) * n + 1)
        P = Projector(n, n, d, np.linspace(0, np.pi, a))
        x = torch.rand(imgs, n, n, device="cuda")
        y = torch.rand(imgs, a, d, device="cuda")
        so = torch.empty(imgs, a, d, device="cuda")
        vo = torch.empty(imgs, n, n, device="cuda")
        reps = max(2, int(2e11 / (n * n * a * imgs)))
        reps = min(reps, 50)
        tf = timeit(lambda: P.fp(x, out=so), reps)
        tb = timeit(lambda: P.bp(y, out=vo), reps)
        U = n * n * a * imgs
        print(json.dumps({"N": n, "A": a, "images": imgs, "fp_ms": round(tf * 1e3, 3), "bp_ms": round(tb * 1e3, 3),
                          "fp_gups": round(U / tf / 1e9, 1), "bp_gups": round(U / tb / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
