#!/usr/bin/env python
"""Measured distances behind the tolerances of the golden-vector tests (tests/test_gpu_parity.py,
tests/test_gpu_solvers.py): how far the device results are from the outputs of the reference's own modules
(tests/golden/*.npz).  Writes profiles/r2_test_margins.json.  Needs a B200."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    import torch
    from whitomo_b200.projector import Projector, Spectrum
    from whitomo_b200 import solvers
    out = {}
    z = np.load(os.path.join(ROOT, "tests", "golden", "white_small.npz"))
    for tag in ("a", "b"):
        n, na = int(z[tag + "_n"]), int(z[tag + "_n_angles"])
        grid, ea, pix = z[tag + "_grid"], z[tag + "_ea"], float(z[tag + "_pixel_size"])
        P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
        sp = Spectrum(z[tag + "_source"] / np.sum(z[tag + "_source"] * grid[:, 1]), grid[:, 1], ea, pix)
        gt = torch.from_numpy(z[tag + "_gt"].astype(np.float32)).cuda()
        c = torch.from_numpy(z[tag + "_c"].astype(np.float32)).cuda()
        meas, _, _ = P.white_fp(sp, gt, want_qmu=False, want_loss=False)
        inten, qmu, loss = P.white_fp(sp, c, meas=meas)
        q = z[tag + "_sinogram"] - z[tag + "_I"]
        # q * mu with the measurement the reference used (its float64 sinogram) instead of the device's own
        meas_ref = torch.from_numpy(z[tag + "_sinogram"].reshape(P.sino_shape).astype(np.float32)).cuda()
        _, qmu_r, _ = P.white_fp(sp, c, meas=meas_ref)
        out["white_" + tag] = {
            "E": int(ea.shape[1]), "n": n,
            "sinogram": rel_l2(meas.cpu().numpy().ravel(), z[tag + "_sinogram"]),
            "I": rel_l2(inten.cpu().numpy().ravel(), z[tag + "_I"]),
            "loss": abs(loss.item() - float(z[tag + "_func"])) / float(z[tag + "_func"]),
            "q_mu_own_measurement": rel_l2(qmu[0].cpu().numpy().reshape(2, -1), q * z[tag + "_mu"]),
            "q_mu_reference_measurement": rel_l2(qmu_r[0].cpu().numpy().reshape(2, -1), q * z[tag + "_mu"]),
            "grad_own_measurement": rel_l2(-P.bp(qmu[0]).cpu().numpy(), z[tag + "_grad"]),
            "grad_reference_measurement": rel_l2(-P.bp(qmu_r[0]).cpu().numpy(), z[tag + "_grad"]),
            "q_over_I": float(np.linalg.norm(q) / np.linalg.norm(z[tag + "_I"])),
        }
    z = np.load(os.path.join(ROOT, "tests", "golden", "barrier_mar.npz"))
    n, na = int(z["n"]), int(z["n_angles"])
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    nb = int(z["n_biter"])
    x = solvers.barrier_least_squares(P, z["projections"].copy(), float(z["i0"]), float(z["bound"]), n_biter=nb)
    x2 = solvers.barrier_least_squares(P, z["projections"].copy(), float(z["i0"]), float(z["bound"]), n_iter=100,
                                       n_biter=nb, t0=0.1, max_iter=1000, do_alpha_decay=True, with_sino_barrier=False)
    out["mar_golden"] = {"run1_with_sinogram_barrier": rel_l2(x.cpu().numpy(), z["ans"]),
                         "run2_without": rel_l2(x2.cpu().numpy(), z["ans_no_ineq"]), "n_biter": nb}
    z = np.load(os.path.join(ROOT, "tests", "golden", "barrier_white.npz"))
    n, na = int(z["n"]), int(z["n_angles"])
    grid = z["grid"]
    P = Projector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
    sp = Spectrum(z["source"] / np.sum(z["source"] * grid[:, 1]), grid[:, 1], z["ea"], float(z["pixel_size"]))
    gt = torch.from_numpy(z["gt"].astype(np.float32)).cuda()
    meas, _, _ = P.white_fp(sp, gt, want_qmu=False, want_loss=False)
    p = {k[2:]: z[k].item() for k in z.files if k.startswith("p_")}
    xx, _ = solvers.white_barrier_stack(P, sp, meas[None] if meas.dim() == 2 else meas, torch.from_numpy(z["x0"].astype(np.float32)).cuda()[None],
                                        ir_beta=float(z["ir_beta"]), **p)
    out["barrier_white_golden"] = {"final_iterate": rel_l2(xx[0].cpu().numpy(), z["ans"])}
    print(json.dumps(out, indent=1))
    with open(os.path.join(ROOT, "gpurun_out", "test_margins.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
