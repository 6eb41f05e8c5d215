#!/usr/bin/env python
"""Per-angle distance of K2 (positions="astra") to the float32 oracle BP: one angle at a time, uniform random sinogram.
Diagnostic for the warped closed form of the reference's running sum (common.cuh: WarpAngle)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import oracle
    from whitomo_b200.projector import Projector
    n, na = int(sys.argv[1]), int(sys.argv[2])
    idx = [int(v) for v in sys.argv[3].split(",")]
    d = int(np.sqrt(2) * n + 1)
    full = np.linspace(0, np.pi, na)
    out = []
    for a in idx:
        ang = full[a:a + 1]
        g = oracle.Geometry(n, n, d, ang)
        y = np.random.default_rng(a).random(g.sino_shape).astype(np.float32)
        ref = oracle.bp(g, y)
        P = Projector(n, n, d, ang, positions="astra")
        v = P.bp(torch.from_numpy(y).cuda()).cpu().numpy()
        diff = (v.astype(np.float64) - ref)
        rel = float(np.linalg.norm(diff) / np.linalg.norm(ref))
        # where is the error: per 256-block rms map (coarse), report the worst block
        blk = 512
        e2 = (diff ** 2).reshape(n // blk, blk, n // blk, blk).sum(axis=(1, 3))
        r2 = (ref.astype(np.float64) ** 2).reshape(n // blk, blk, n // blk, blk).sum(axis=(1, 3))
        worst = np.unravel_index(np.argmax(e2), e2.shape)
        out.append({"angle": a, "theta_deg": float(np.degrees(full[a])), "rel_l2": rel,
                    "worst_block": [int(worst[0]), int(worst[1])], "worst_block_rel": float(np.sqrt(e2[worst] / max(r2[worst], 1e-300)))})
        print(out[-1], flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "bp_angle_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
