"""Multi-GPU decomposition of the hot path.  One process per GPU, torch.distributed for plumbing.

* Slice stacks (C4) shard naturally: contiguous slice ranges per rank, NO data-path collective
  (``shard_range``; an optional scalar all-reduce of the loss is for logging only).
* A single oversized slice (C5: 8192^2, 4096 angles) is split by ANGLE RANGE: every rank keeps the full volume
  and owns the sinogram rows of its angles.  FP, the white-beam residual and mu are local to a ray, so they
  need no communication; each rank backprojects its own angle range and the partial volumes are summed with
  one all-reduce per iteration (NCCL over NVLink on GPUs; gloo in the CPU tests).

``proj`` is any object with the Projector interface (fp / bp / white_fp taking ``angle_range``), so the
decomposition logic is testable on CPU with a stand-in backed by the oracle.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """contiguous [start, stop) of `n_items` for `rank`; sizes differ by at most one"""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def angle_ranges(n_angles, world, align=8):
    """angle ranges per rank, split points aligned to the forward projector's 8-angle tiles where possible"""
    cuts = [0]
    for r in range(1, world):
        c = n_angles * r // world
        if n_angles >= 8 * align * world:
            c = int(round(c / float(align))) * align
        cuts.append(min(max(c, cuts[-1]), n_angles))
    cuts.append(n_angles)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def angle_ranges_by_cost(tile_first, tile_count, n_angles, world, fp_share=0.55, tile_overhead=0.12, tile_angles=8):
    """Angle ranges per rank that balance the projector COST instead of the angle count.  The forward projector runs
    one CTA column per angle tile; a tile that holds fewer than 8 angles (tiles around 45/135 degrees: their rays
    diverge faster than one staging window allows) gathers proportionally less but keeps its per-tile overhead
    (staging, idle ray slots): measured, an angle of such a tile costs K1 about 8 % more (profiles/r1_c5_anglesplit.json;
    weighting tiles as if their cost did not depend on their angle count made the split WORSE: profiles/r2_bench_8gpu.json
    first run).  The backprojector's cost is per angle.  Cuts fall on tile boundaries.  tile_first / tile_count: the plan
    of wt_geom_plan."""
    cost = [fp_share * (tile_overhead * tile_angles + (1.0 - tile_overhead) * c) + (1.0 - fp_share) * c for c in tile_count]
    total = float(sum(cost))
    cuts, acc, r = [0], 0.0, 1
    for f, c, w in zip(tile_first, tile_count, cost):
        # cut before this tile if that brings the running cost closer to the rank's share
        while r < world and acc + 0.5 * w >= total * r / world:
            cuts.append(f)
            r += 1
        acc += w
    while len(cuts) < world:
        cuts.append(n_angles)
    cuts.append(n_angles)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def plan_tiles(proj):
    """(tile_first, tile_count) of the forward projector's angle tiles (host-only diagnostic wt_geom_plan)"""
    import ctypes
    from . import _capi
    lib = _capi.load()
    ang = proj.angles
    mx = proj.nangles
    first, count = (ctypes.c_int * mx)(), (ctypes.c_int * mx)()
    nt = ctypes.c_int()
    _capi.check(lib.wt_geom_plan(proj.rows, proj.cols, proj.ndet, proj.det_width,
                                 ang.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), proj.nangles, mx,
                                 first, count, None, None, ctypes.byref(nt), None, None))
    return list(first[:nt.value]), list(count[:nt.value])


class AngleSplit(object):
    """W and W^T of one slice distributed over the ranks of `group` by angle range.

    balance="cost" (default) cuts the angle list where the projector cost is equal (angle_ranges_by_cost),
    "count" where the angle count is.  `bp` overlaps the sum of the partial backprojections with their computation:
    the volume is backprojected in `bp_blocks` row blocks and every finished block goes into an asynchronous
    all-reduce while the next block is computed; only the last block's reduction is exposed."""

    def __init__(self, proj, group=None, balance="cost", bp_blocks=8):
        self.proj = proj
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bp_blocks = int(bp_blocks)
        if balance == "cost" and self.world > 1 and hasattr(proj, "handle"):
            first, count = plan_tiles(proj)
            self.ranges = angle_ranges_by_cost(first, count, proj.nangles, self.world)
        else:
            self.ranges = angle_ranges(proj.nangles, self.world)
        self.range = self.ranges[self.rank]

    def _allreduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def fp(self, x):
        """rows of this rank's angles of W x (other rows zero); no communication"""
        return self.proj.fp(x, angle_range=self.range)

    def bp(self, y, scale=1.0):
        """W^T y: local partial backprojection of this rank's angles, summed over the ranks.  Row block by row block
        when the projector offers `bp_rows` (the all-reduce of a finished block runs while the next one is computed),
        otherwise (the CPU stand-ins of the tests) one backprojection and one all-reduce."""
        if self.world == 1 or self.bp_blocks <= 1 or not hasattr(self.proj, "bp_rows"):
            return self._allreduce(self.proj.bp(y, angle_range=self.range, scale=scale))
        import torch
        lead = tuple(y.shape[:-2])
        n = 1
        for v in lead:
            n *= v
        out = torch.empty((n,) + tuple(self.proj.vol_shape), dtype=torch.float32, device=self.proj.device)
        pending = []
        for i, (b0, b1, written) in enumerate(self.proj.bp_row_blocks(n, self.bp_blocks)):
            self.proj.bp_rows(y, out, b0, b1, angle_range=self.range, scale=scale, reuse_packed=i > 0)
            for lo, hi in written:       # row ranges are contiguous only for one image: reduce image by image
                for im in range(n):
                    pending.append(dist.all_reduce(out[im, lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in pending:
            w.wait()
        return out.reshape(lead + tuple(self.proj.vol_shape))

    def white_func_grad(self, spectrum, conc, meas):
        """(loss, grad) of the white-beam data term with rays split by angle: loss and grad are all-reduced"""
        _, qmu, loss = self.proj.white_fp(spectrum, conc, meas=meas, want_intensity=False, angle_range=self.range)
        g = self.proj.bp(qmu.reshape((-1,) + tuple(qmu.shape[-2:])), angle_range=self.range, scale=-1.0)
        return self._allreduce(loss), self._allreduce(g).reshape(conc.shape)

    def gather_sinogram(self, y_local):
        """full sinogram on every rank from the locally owned rows (rows outside the range must be zero)"""
        return self._allreduce(y_local.clone())


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs closest to its GPU (NVML's ideal affinity) so that pinned staging
    buffers are first-touched on the GPU's own NUMA node: with 8 ranks per box every rank otherwise contends for
    one socket's memory controllers during H2D/D2H.  Best effort: returns False if NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(device_index))
        return True
    except Exception:
        return False
