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


class AngleSplit(object):
    """W and W^T of one slice distributed over the ranks of `group` by angle range."""

    def __init__(self, proj, group=None):
        self.proj = proj
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.range = angle_ranges(proj.nangles, self.world)[self.rank]

    def _allreduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def fp(self, x):
        """rows of this rank's angles of W x (other rows zero); no communication"""
        return self.proj.fp(x, angle_range=self.range)

    def bp(self, y, scale=1.0):
        """W^T y: local partial backprojection of this rank's angles, then one all-reduce"""
        return self._allreduce(self.proj.bp(y, angle_range=self.range, scale=scale))

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
