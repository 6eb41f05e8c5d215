"""C5: one oversized slice split by angle range over the ranks (torchrun, NCCL).  Checks the distributed
FP/BP against the single-GPU operators and times one SIRT-like iteration (local FP rows, local partial BP,
one all-reduce of the partial volumes).  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_anglesplit.py [N A]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from whitomo_b200.projector import Projector
from whitomo_b200.distributed import AngleSplit


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    na = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d = int(np.sqrt(2) * n + 1)
    P = Projector(n, n, d, np.linspace(0, np.pi, na), device=dev)
    sp = AngleSplit(P)
    gen = torch.Generator(device=dev).manual_seed(5)           # same data on every rank
    x = torch.rand((n, n), device=dev, generator=gen)
    y = torch.rand((na, d), device=dev, generator=gen)
    # correctness: distributed == single GPU
    full_fp, full_bp = P.fp(x), P.bp(y)
    loc = sp.fp(x)
    a0, a1 = sp.range
    ok_fp = bool(torch.equal(loc[a0:a1], full_fp[a0:a1]))
    dbp = sp.bp(y)
    err_bp = float(torch.linalg.vector_norm((dbp - full_bp).double()) / torch.linalg.vector_norm(full_bp.double()))

    def iteration():
        r = sp.fp(x)                 # own sinogram rows
        return sp.bp(r)              # partial BP + all-reduce

    for _ in range(2):
        iteration()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # where the time goes on each rank: local FP rows, local partial BP, the all-reduce
    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    r_loc = sp.fp(x)
    part = P.bp(r_loc, angle_range=sp.range)
    parts = torch.tensor([timed(lambda: sp.fp(x)), timed(lambda: P.bp(r_loc, angle_range=sp.range)),
                          timed(lambda: sp._allreduce(part))], device=dev, dtype=torch.float64)
    allparts = [torch.zeros_like(parts) for _ in range(world)]
    if world > 1:
        dist.all_gather(allparts, parts)
    else:
        allparts = [parts]
    # single-GPU time of the same iteration for reference
    e0.record()
    for _ in range(reps):
        P.bp(P.fp(x))
    e1.record()
    torch.cuda.synchronize()
    t1 = e0.elapsed_time(e1) / reps
    if rank == 0:
        print(json.dumps({"config": "C5-style angle split", "N": n, "A": na, "ranks": world, "range_rank0": [a0, a1],
                          "fp_rows_bitwise_equal": ok_fp, "bp_rel_l2_vs_single_gpu": err_bp,
                          "ms_per_iteration_split": float(t.item()), "ms_per_iteration_single_gpu": t1,
                          "gups_split": 2.0 * n * n * na / (float(t.item()) * 1e-3) / 1e9,
                          "per_rank_ms_fp_bp_allreduce": [[round(v, 3) for v in q.tolist()] for q in allparts],
                          "angle_ranges": [list(r) for r in __import__("whitomo_b200.distributed", fromlist=["x"]).angle_ranges(na, world)]}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
