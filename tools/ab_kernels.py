#!/usr/bin/env python
"""A/B timing of wt_fp / wt_bp through the bare C ABI of a given build of the library (development aid).
    python tools/ab_kernels.py <lib.so> N,A,images [N,A,images ...]        (WT_POSITIONS selects the position mode)"""
import ctypes
import json
import os
import sys

import numpy as np
import torch


def main():
    lib = ctypes.CDLL(sys.argv[1])
    vp, ci, cf, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    lib.wt_geom_create.argtypes = [ci, ci, ci, cf, ctypes.POINTER(cf), ci, ctypes.POINTER(vp)]
    lib.wt_workspace_bytes.restype = sz
    lib.wt_workspace_bytes.argtypes = [vp, ci]
    lib.wt_fp.argtypes = [vp, vp, vp, ci, ci, ci, vp, sz, vp]
    lib.wt_bp.argtypes = [vp, vp, vp, ci, ci, ci, cf, vp, sz, vp]
    for arg in sys.argv[2:]:
        n, a, imgs = (int(v) for v in arg.split(","))
        d = int(np.sqrt(2) * n + 1)
        ang = np.linspace(0, np.pi, a).astype(np.float32)
        h = vp()
        assert lib.wt_geom_create(n, n, d, 1.0, ang.ctypes.data_as(ctypes.POINTER(cf)), a, ctypes.byref(h)) == 0
        x = torch.rand(imgs, n, n, device="cuda")
        y = torch.rand(imgs, a, d, device="cuda")
        so, vo = torch.empty_like(y), torch.empty_like(x)
        nb = lib.wt_workspace_bytes(h, imgs)
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        st = vp(torch.cuda.current_stream().cuda_stream)
        fp = lambda: lib.wt_fp(h, vp(x.data_ptr()), vp(so.data_ptr()), imgs, 0, a, vp(ws.data_ptr()), nb, st)
        bp = lambda: lib.wt_bp(h, vp(y.data_ptr()), vp(vo.data_ptr()), imgs, 0, a, 1.0, vp(ws.data_ptr()), nb, st)
        reps = min(20, max(2, int(1e11 / (n * n * a * imgs))))
        res = {"lib": os.path.basename(sys.argv[1]), "positions": os.environ.get("WT_POSITIONS", "default"), "N": n, "A": a, "images": imgs}
        for name, fn in (("fp", fp), ("bp", bp)):
            assert fn() == 0
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / reps
            res[name + "_ms"] = round(t, 3)
            res[name + "_gups"] = round(n * n * a * imgs / t / 1e6, 1)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
