#!/usr/bin/env python
"""A/B of two builds of the library, one process each (development aid): wt_fp / wt_bp through the bare C ABI, timed with
CUDA events, outputs compared (max |a-b| relative to max |a|, and whether the bits agree).
    python tools/ab_compare.py <old.so> <new.so> N,A,images[,positions] ...      positions: astra (default) | exact"""
import ctypes
import json
import sys

import numpy as np
import torch

vp, ci, cf, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


def load(path):
    lib = ctypes.CDLL(path)
    lib.wt_geom_create_ex.argtypes = [ci, ci, ci, cf, ctypes.POINTER(cf), ci, ci, ctypes.POINTER(vp)]
    lib.wt_workspace_bytes.restype = sz
    lib.wt_workspace_bytes.argtypes = [vp, ci]
    lib.wt_fp.argtypes = [vp, vp, vp, ci, ci, ci, vp, sz, vp]
    lib.wt_bp.argtypes = [vp, vp, vp, ci, ci, ci, cf, vp, sz, vp]
    return lib


def run(lib, n, a, imgs, positions, x, y):
    d = y.shape[-1]
    ang = np.linspace(0, np.pi, a).astype(np.float32)
    h = vp()
    rc = lib.wt_geom_create_ex(n, n, d, 1.0, ang.ctypes.data_as(ctypes.POINTER(cf)), a, positions, ctypes.byref(h))
    assert rc == 0, rc
    so, vo = torch.empty_like(y), torch.empty_like(x)
    nb = lib.wt_workspace_bytes(h, imgs)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    st = vp(torch.cuda.current_stream().cuda_stream)
    fp = lambda: lib.wt_fp(h, vp(x.data_ptr()), vp(so.data_ptr()), imgs, 0, a, vp(ws.data_ptr()), nb, st)
    bp = lambda: lib.wt_bp(h, vp(y.data_ptr()), vp(vo.data_ptr()), imgs, 0, a, 1.0, vp(ws.data_ptr()), nb, st)
    reps = min(50, max(3, int(3e10 / (n * n * a * imgs))))
    res = {}
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
        res[name + "_ms"] = round(e0.elapsed_time(e1) / reps, 4)
    return res, so.clone(), vo.clone()


def child(path, args, out):
    """one library per process: template statics (GNU unique symbols) are shared between two builds loaded side by side"""
    lib = load(path)
    results = []
    for k, arg in enumerate(args):
        parts = arg.split(",")
        n, a, imgs = (int(v) for v in parts[:3])
        positions = 1 if len(parts) > 3 and parts[3] == "exact" else 0
        d = int(np.sqrt(2) * n + 1)
        torch.manual_seed(1)
        x = torch.rand(imgs, n, n, device="cuda")
        y = torch.rand(imgs, a, d, device="cuda")
        r, so, vo = run(lib, n, a, imgs, positions, x, y)
        torch.save({"so": so.cpu(), "vo": vo.cpu()}, "%s.%d.pt" % (out, k))
        results.append(r)
    json.dump(results, open(out + ".json", "w"))


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[4:], sys.argv[3])
    import subprocess
    import tempfile
    tmp = tempfile.mkdtemp()
    args = sys.argv[3:]
    for tag, path in (("old", sys.argv[1]), ("new", sys.argv[2])):
        subprocess.check_call([sys.executable, __file__, "--child", path, "%s/%s" % (tmp, tag)] + args)
    ro, rn = (json.load(open("%s/%s.json" % (tmp, t))) for t in ("old", "new"))
    for k, arg in enumerate(args):
        parts = arg.split(",")
        out = {"N": int(parts[0]), "A": int(parts[1]), "images": int(parts[2]), "positions": parts[3] if len(parts) > 3 else "astra"}
        to, tn = (torch.load("%s/%s.%d.pt" % (tmp, t, k)) for t in ("old", "new"))
        for key, name in (("fp", "so"), ("bp", "vo")):
            out[key + "_ms_old"], out[key + "_ms_new"] = ro[k][key + "_ms"], rn[k][key + "_ms"]
            out[key + "_speedup"] = round(ro[k][key + "_ms"] / rn[k][key + "_ms"], 3)
            out[key + "_maxrel"] = float((to[name] - tn[name]).abs().max() / to[name].abs().max())
            out[key + "_bits_equal"] = bool(torch.equal(to[name], tn[name]))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
