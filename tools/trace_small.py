"""Development aid: where the CTAs of K1 / K2 spend their time on a small problem.  Needs a -DWT_TRACE=1 build:
  nvcc ... -DWT_TRACE=1 -o tools/bin/libwt_trace.so whitomo_b200/csrc/api.cu;  WT_LIB=tools/bin/libwt_trace.so python tools/trace_small.py [N A]
Time stamps per CTA (thread 0): 0 entry (globaltimer ns), 1 entry, 2 set-up done, 3 first stage landed, 4 main loop done,
5 cluster reduction done, 6 epilogue done (clock64), 7 exit (globaltimer ns)."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from whitomo_b200.projector import Projector
from whitomo_b200 import _capi

n, a = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 180)
d = int(np.sqrt(2) * n + 1)
P = Projector(n, n, d, np.linspace(0, np.pi, a))
y = P.fp(torch.rand(1, n, n, device="cuda"))
P.sirt(y, 20)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["WT_LIB"])
lib.wt_debug_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
for k, name in ((0, "K1 fp_kernel"), (1, "K2 bp_kernel")):
    buf = np.zeros((2048, 8), dtype=np.uint64)
    assert lib.wt_debug_trace(k, buf.ctypes.data, 2048) == 0
    used = buf[:, 0] > 0
    t = buf[used].astype(np.int64)
    start = t[:, 0] - t[:, 0].min()
    end = (t[:, 7] - t[:, 0].min())[t[:, 7] > 0]
    cyc = t[:, 1:7] - t[:, 1:2]
    ok = (t[:, 6] > 0)
    rep = {"kernel": name, "ctas_traced": int(used.sum()), "cta_start_ns": [int(start.min()), int(np.median(start)), int(start.max())],
           "cta_end_ns": [int(end.min()), int(np.median(end)), int(end.max())]}
    names = ["setup_done", "first_stage_landed", "main_loop_done", "cluster_reduce_done", "epilogue_done"]
    for i, nm in enumerate(names, start=1):
        col = cyc[:, i]
        col = col[(col > 0) & (col < 10**9)]
        if len(col): rep[nm + "_cycles_from_entry(min,med,max)"] = [int(col.min()), int(np.median(col)), int(col.max())]
    print(json.dumps(rep))
