#!/usr/bin/env python
"""Distance of K1/K2 to the reference's arithmetic at the BASELINE configurations, on FULL angle sets.

For C3 (1024^2 x 720), C4 (2048^2 x 1800, one slice) and C5 (8192^2, every 32nd of its 4096 angles) one image goes
through the CUDA projectors (both position modes) and through the CPU oracle on all host threads: the float32
restatement of ASTRA's running-sum projector (oracle.fp / oracle.bp, what the reference computes) and the float64
closed-form arbiter (oracle.fp64 / oracle.bp64).  Three relative-L2 distances per operator and input:

    gpu_vs_oracle32    the parity number (north_star: <= 1e-4)
    gpu_vs_f64         distance of the kernel to exact arithmetic
    oracle32_vs_f64    what the reference's float32 running sum costs it

Inputs: FP of the two-disc phantom; BP of that phantom's sinogram ("smooth") and of a uniform random sinogram.
Writes profiles/r2_parity_fullsize.json.  Needs a B200 (the oracle legs take 1-2 minutes of host time).

    python tools/parity_fullsize.py [--configs C3,C4,C5] [--out profiles/r2_parity_fullsize.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    "C2": dict(n=512, n_angles=360, stride=1),
    "C3": dict(n=1024, n_angles=720, stride=1),
    "C4": dict(n=2048, n_angles=1800, stride=1),
    "C5": dict(n=8192, n_angles=4096, stride=32),
}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_config(name, modes=("astra", "exact"), with_f64=True):
    """-> dict of distances for one configuration (see module docstring)"""
    import torch
    import oracle
    from whitomo_b200 import phantoms
    from whitomo_b200.projector import Projector
    cfg = CONFIGS[name]
    n, na = cfg["n"], cfg["n_angles"]
    ang = np.linspace(0, np.pi, na)[::cfg["stride"]]
    d = int(np.sqrt(2) * n + 1)
    g = oracle.Geometry(n, n, d, ang)
    discs = phantoms.two_discs(n)
    x = (discs[0] + 0.5 * discs[1]).astype(np.float32)
    t0 = time.perf_counter()
    p32 = oracle.fp(g, x, threads=0)
    y_rand = np.random.default_rng(1).random(g.sino_shape).astype(np.float32)
    b32 = {"smooth": oracle.bp(g, p32, threads=0), "random": oracle.bp(g, y_rand, threads=0)}
    t_o32 = time.perf_counter() - t0
    p64 = b64 = None
    t_f64 = 0.0
    if with_f64:
        t0 = time.perf_counter()
        p64 = oracle.fp64(g, x, threads=0)
        b64 = {"smooth": oracle.bp64(g, p32, threads=0), "random": oracle.bp64(g, y_rand, threads=0)}
        t_f64 = time.perf_counter() - t0
    out = {"n": n, "n_angles_config": na, "n_angles_run": int(len(ang)), "ndet": d,
           "oracle32_seconds": round(t_o32, 1), "f64_seconds": round(t_f64, 1), "host_threads": oracle.num_threads()}
    if with_f64:
        out["oracle32_vs_f64"] = {"fp": rel_l2(p32, p64), "bp_smooth": rel_l2(b32["smooth"], b64["smooth"]),
                                  "bp_random": rel_l2(b32["random"], b64["random"])}
    for mode in modes:
        P = Projector(n, n, d, ang, positions=mode)
        xs = torch.from_numpy(x).cuda()
        res = {}
        # the image alone (1-image kernels) and inside a batch of 4 (float4 kernels): both must match
        for tag, batch in (("single", 1), ("batch4", 4)):
            xb = xs.unsqueeze(0).repeat(batch, 1, 1)
            p = P.fp(xb)[batch - 1].cpu().numpy()
            r = {"fp": {"gpu_vs_oracle32": rel_l2(p, p32)}}
            if with_f64:
                r["fp"]["gpu_vs_f64"] = rel_l2(p, p64)
            for kind, y in (("smooth", p32), ("random", y_rand)):
                yb = torch.from_numpy(y).cuda().unsqueeze(0).repeat(batch, 1, 1)
                v = P.bp(yb)[batch - 1].cpu().numpy()
                r["bp_" + kind] = {"gpu_vs_oracle32": rel_l2(v, b32[kind])}
                if with_f64:
                    r["bp_" + kind]["gpu_vs_f64"] = rel_l2(v, b64[kind])
            res[tag] = r
        out[mode] = res
        del P
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C3,C4,C5")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_parity_fullsize.json"))
    ap.add_argument("--no-f64", action="store_true")
    args = ap.parse_args()
    res = {}
    for name in args.configs.split(","):
        res[name] = run_config(name, with_f64=not args.no_f64)
        print(name, json.dumps(res[name]), flush=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
