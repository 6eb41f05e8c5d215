"""Device-resident solver loops for the reference's concrete scenarios.  Everything between the input upload and
the result download stays in HBM; each inner iteration is K1 (fused epilogue) -> K2 -> K3, i.e. K FP + K BP
(white-beam) or 1-2 FP + 1 BP (mono) where the reference spends 3K FP + K BP / >= 4 FP + 2 BP on the same
arithmetic.

  white_barrier_stack     src/barrier_method_gogo.py:79-94,179-193 driven by src/barrier_method.py:155-240,
                          goal = WhiteProjection.func/grad, over a stack of independent slices
  white_rec_stack         src/white_rec.py:199-244,338-368 (plain gradient descent `Iteration`)
  barrier_least_squares   teeth_recon/experiments.py:651-749 (mono metal-artifact reduction, HoughBarrier)
  sirt                    src/astra_proxy.py:69-84 / teeth_recon/experiments.py:352-356

Slices of a stack are independent problems sharing one geometry; the early-stop rule of barrier_method
(src/barrier_method.py:208-229) is therefore applied per slice through an `active` mask.
"""
import math

import numpy as np
import torch

from . import _capi


def _as_stack(x, K):
    """(K,N,N) or (S,K,N,N) -> (S,K,N,N) float32 contiguous clone"""
    if x.dim() == 3:
        x = x.unsqueeze(0)
    assert x.dim() == 4 and x.shape[1] == K
    return x.to(torch.float32).contiguous().clone()


def white_barrier_stack(proj, spectrum, meas, x0, n_iter=50, n_biter=20, t0=0.1, t_step=0.5, alpha=1.0,
                        beta_reg=1.0, ir_beta=0.05, max_iter=5000, do_alpha_decay=True, min_delta_loss=1e-3,
                        n_steps_without_progress=15, add_stat_cb=None, max_total_iters=None, compact=True, stat_cb=None,
                        fused=True):
    """Log-barrier reconstruction of element concentrations 0 <= c <= 1 with the 0.05*|c0 c1| regulariser.

    Defaults are the reference's run (src/barrier_method_gogo.py:184-192).  ``meas`` (S, A, D) intensities,
    ``x0`` (S, K, N, N) feasible start.  Returns (x, info) with info = dict(iters, loss (S,), trace list,
    slice_iters = projector iterations actually spent per slice).
    Per inner iteration and slice: one fused white-beam FP (K projections) and K BP whose epilogue takes the step
    (``fused``, three launches per iteration; with a statistics callback, K > 2 or ``fused=False``: the separate entry
    points wt_white_fp / wt_bp / wt_barrier_update, seven launches, same arithmetic).
    The loss of barrier_method's early-stop test at the new point is f(x') + penalty(x'); f(x') is the by-product
    of the FP that the next iteration's gradient needs anyway, so no projection is spent on statistics.
    ``add_stat_cb`` (alias ``stat_cb``) receives what the reference's callback receives (src/barrier_method.py:182-199):
    the tuple ``(goal_stat, reg_stats, bf_stats)`` of ``StepStat(x, func, grad)`` -- goal: f(x) and the TOTAL gradient
    (the reference accumulates into the goal gradient in place); ``reg_stats['ineq_reg']``: the
    0.05 |c0 c1| term and its gradient; ``bf_stats['conc_non_negative' | 'conc_less_than_one']``: phi and grad phi of
    the two box barriers -- as device tensors with a leading axis over the slices of the working set (their stack
    indices are ``goal_stat.x.slices``), at the reference's cadence (every inner iteration while the outer iteration
    index is < 30 or a multiple of 10, B6).  f(x) comes from the projector's epilogue; the other statistics cost a few
    element-wise passes, spent only when a callback is installed.
    Early stop is per slice (each slice is its own barrier_method problem).  With ``compact`` the slices that left
    a barrier level early are taken out of the working set until the next level, so no FP/BP is spent on them
    (the working set never drops below 3 images, which keeps every slice on the batched kernels and its result
    independent of what the other slices do).
    """
    K = spectrum.K
    x = _as_stack(x0, K)
    S = x.shape[0]
    meas = meas.reshape((S,) + proj.sino_shape).to(torch.float32).contiguous()
    dev = x.device
    add_stat_cb = add_stat_cb if add_stat_cb is not None else stat_cb
    if not bool(((x > 0) & (x < 1)).all()):         # src/barrier_method.py:128-152 (box projection, then re-check)
        x.clamp_(0, 1)
        if not bool(((x >= 0) & (x <= 1)).all()):
            return x, None
    t = float(t0)
    total = 0
    trace = []
    slice_iters = torch.zeros(S, dtype=torch.int64, device=dev)
    min_work = -(-3 // K)                            # slices that keep the working set at >= 3 images
    want_pen = n_steps_without_progress is not None
    # Fused iteration (K2 with the update in its epilogue -> K1 -> one reduction: three launches, no gradient / q*mu
    # arrays) whenever nobody asks for the gradient (statistics callback) and the kernels offer it; else the separate
    # entry points (seven launches).  Same arithmetic either way.
    fused = fused and add_stat_cb is None and hasattr(proj, "step_supported") and proj.step_supported(spectrum, min(S, min_work))
    qmu = None
    if fused:
        state = proj.new_step_state(spectrum, S)
        f_loss, _ = proj.white_barrier_step(spectrum, x, meas, state, init=True)
        state_ok = True                              # `state` holds the packed copies of the full stack x
    else:
        _, qmu, f_loss = proj.white_fp(spectrum, x, meas=meas, want_intensity=False)
    done = False
    for i_biter in range(n_biter):
        # working set of this barrier level: all slices; shrinks as slices stop early
        idx = torch.arange(S, device=dev)
        xw, mw, qw, fw = x, meas, qmu, f_loss
        if fused:
            st_w = state
            if not state_ok:                         # the last level parked compacted results into x: re-pack
                f_loss, _ = proj.white_barrier_step(spectrum, x, meas, state, init=True)
                fw, state_ok = f_loss, True
        prev = None
        stall = torch.zeros(S, dtype=torch.int32, device=dev)
        active = torch.ones(S, dtype=torch.bool, device=dev)
        for i_iter in range(n_iter):
            n_w = xw.shape[0]
            if fused:
                fw, pen = proj.white_barrier_step(spectrum, xw, mw, st_w, alpha=alpha, beta_reg=beta_reg, t=t, reg_w=ir_beta,
                                                  active=active, want_penalty=want_pen)
            else:
                g = proj.bp(qw.reshape((n_w * K,) + proj.sino_shape), scale=-1.0)
                if add_stat_cb is not None and (i_biter % 10 == 0 or i_biter < 30):
                    add_stat_cb(_white_barrier_stats(xw, idx, fw, g.reshape(xw.shape), beta_reg, ir_beta, t))
                pen = proj.barrier_update(xw, g, alpha, beta_reg, t, reg_w=ir_beta, active=active, want_penalty=want_pen)
                _, qw, fw = proj.white_fp(spectrum, xw, meas=mw, want_intensity=False)
            slice_iters[idx] += active.to(torch.int64)
            total += 1
            if n_steps_without_progress is not None:
                loss = fw + pen
                if prev is not None:
                    progress = (prev - loss) >= min_delta_loss       # False for nan/inf, like the reference
                    stall = torch.where(active, torch.where(progress, torch.zeros_like(stall), stall + 1), stall)
                prev = loss if prev is None else torch.where(active, loss, prev)
                active = active & (stall < n_steps_without_progress)
                trace.append((idx, loss.detach()))
                n_active = int(active.sum().item())
                if n_active == 0:
                    break
                if compact and n_active < n_w and n_active >= min_work:
                    if xw is not x:
                        x[idx], f_loss[idx] = xw, fw                      # park what leaves the working set
                        if not fused:
                            qmu[idx] = qw
                    else:
                        qmu, f_loss = qw, fw
                    keep = active.nonzero(as_tuple=True)[0]
                    idx = idx[keep]
                    xw, mw, fw = x[idx].contiguous(), meas[idx].contiguous(), fw[keep]
                    prev, stall, active = prev[keep], stall[keep], active[keep]
                    if fused:                        # a new working set: its own packed state, forward model redone
                        st_w = proj.new_step_state(spectrum, xw.shape[0])
                        fw, _ = proj.white_barrier_step(spectrum, xw, mw, st_w, init=True)
                        state_ok = False
                    else:
                        qw = qw[keep].contiguous()
            if max_total_iters is not None and total >= max_total_iters:
                done = True
                break
        if xw is not x:
            x[idx], f_loss[idx] = xw, fw
            if not fused:
                qmu[idx] = qw
        else:
            qmu, f_loss = qw, fw
        t *= t_step
        if do_alpha_decay:
            alpha = alpha * math.sqrt(t_step)
            n_iter = min(int(n_iter / t_step), max_iter)
        if done:
            break
    return x, dict(iters=total, loss=f_loss, trace=trace, slice_iters=slice_iters)


class _SnapShot(torch.Tensor):
    """a tensor that remembers which slices of the stack it holds (StepStat.x of the fused loops)"""


def _white_barrier_stats(xw, idx, f_loss, g, beta_reg, ir_beta, t):
    """the reference's (goal_stat, reg_stats, bf_stats) for the gogo problem (src/barrier_method.py:182-199 with the
    tuples of src/barrier_method_gogo.py:79-94), batched over the working set"""
    from .barrier_method import StepStat
    snap = xw.clone().as_subclass(_SnapShot)
    snap.slices = idx.clone()
    K = xw.shape[1]
    inf = torch.full_like(xw, float("inf"))
    reg_val = ir_beta * (xw[:, 0] * xw[:, 1]).abs() if K == 2 else torch.zeros_like(xw[:, 0])
    reg_grad = torch.zeros_like(xw)
    if K == 2:
        sgn = torch.sign(xw[:, 0] * xw[:, 1])
        reg_grad[:, 0], reg_grad[:, 1] = ir_beta * xw[:, 1] * sgn, ir_beta * xw[:, 0] * sgn
    phi_lo = torch.where(xw > 0, -torch.log(xw), inf)              # f = -x:     phi = -log(x),   grad phi = -1/x
    phi_hi = torch.where(xw < 1, -torch.log1p(-xw), inf)           # f = x - 1:  phi = -log(1-x), grad phi = 1/(1-x)
    g_lo, g_hi = -1.0 / xw, 1.0 / (1.0 - xw)
    total = g + beta_reg * reg_grad + beta_reg * t * (g_lo + g_hi)
    return (StepStat(snap, f_loss.clone(), total),
            {"ineq_reg": StepStat(snap, reg_val, reg_grad)},
            {"conc_non_negative": StepStat(snap, phi_lo, g_lo), "conc_less_than_one": StepStat(snap, phi_hi, g_hi)})


def white_rec_stack(proj, spectrum, meas, c0, n_iters=1000, alpha=0.7, beta_reg=1e-2, clip_concentrations=True,
                    dissimilarity_constraint=True, triv_conc_constraint=True, alternate=True, start_iter=0, fused=True):
    """Plain gradient descent of src/white_rec.py:199-244,338-368 (flags as at :342-345, alpha/beta at :42-43):
    step = alpha*(grad + beta*(d|c0 c1| + c(c(4c-3)+1))); iteration i updates element 0 when i % 4 == 0 else
    element 1... (the reference zeroes step[1] when i % 4 == 0 and step[0] otherwise, :229-233); clip to [0, 1].
    ``start_iter``: iteration number of the first step (a run continued from an earlier call keeps the alternation).
    Returns (c, losses) with losses[i] = ||q_i|| + ||beta c0 c1|| per slice (:236-237)."""
    K = spectrum.K
    c = _as_stack(c0, K)
    S = c.shape[0]
    meas = meas.reshape((S,) + proj.sino_shape).to(torch.float32).contiguous()
    losses = []
    mode = _capi.UPD_STEP_CLIP if clip_concentrations else _capi.UPD_STEP_ONLY
    reg_w = 1.0 if (dissimilarity_constraint and K == 2) else 0.0
    triv_w = 1.0 if triv_conc_constraint else 0.0
    fused = fused and hasattr(proj, "step_supported") and proj.step_supported(spectrum, S)
    if fused:       # three launches per iteration (K2 + step, K1, reduction), see white_barrier_stack
        state = proj.new_step_state(spectrum, S)
        f_loss, _ = proj.white_barrier_step(spectrum, c, meas, state, init=True)
    f_hist = []                         # per iteration: f (S,) float64; combined with ||c0 c1|| after the loop
    want_r = dissimilarity_constraint and K == 2
    pair_norm2 = getattr(proj, "pair_norm2", None)        # (the numpy drop-in of the tests has none)
    r_part = torch.empty((n_iters, S, _capi.PAIR_NORM_BLOCKS), dtype=torch.float64, device=c.device) \
        if (want_r and pair_norm2 is not None and n_iters > 0) else None
    r_hist = []
    for it in range(start_iter, start_iter + n_iters):
        if not fused:
            _, qmu, f_loss = proj.white_fp(spectrum, c, meas=meas, want_intensity=False)
            g = proj.bp(qmu.reshape((S * K,) + proj.sino_shape), scale=-1.0)
        f_hist.append(f_loss)
        if r_part is not None:          # sum (c0 c1)^2 as partial sums: one launch, reduced after the loop
            pair_norm2(c, out=r_part[it - start_iter])
        elif want_r:
            r_hist.append(torch.linalg.vector_norm((c[:, 0] * c[:, 1]).reshape(S, -1), dim=1))
        mask = 0xFFFFFFFF
        if alternate and K == 2:
            mask = 0b01 if it % 4 == 0 else 0b10
        if fused:
            f_loss, _ = proj.white_barrier_step(spectrum, c, meas, state, alpha=alpha, beta_reg=beta_reg, t=0.0, reg_w=reg_w,
                                                triv_w=triv_w, mode=mode, elem_mask=mask)
        else:
            proj.barrier_update(c, g, alpha, beta_reg, 0.0, reg_w=reg_w, triv_w=triv_w, mode=mode, elem_mask=mask)
    # losses[i] = ||q_i|| + ||beta c0 c1||: one small kernel per iteration inside the loop, the rest once here
    if f_hist:
        total = torch.sqrt(2.0 * torch.stack(f_hist))
        if r_part is not None:
            total = total + abs(beta_reg) * torch.sqrt(r_part.sum(-1))
        elif r_hist:
            total = total + abs(beta_reg) * torch.stack(r_hist).double()
        losses = list(total.unbind(0))
    return c, losses


def sirt(proj, sino, n_iters=100):
    return proj.sirt(sino, n_iters)


def ray_transform_from_projection(p, bound, i0):
    """teeth_recon/experiments.py:290-293: m = p <= bound; p[m] = bound; r = log i0 - log p (p is modified, like
    the reference).  Torch tensors or numpy arrays."""
    m = p <= bound
    p[m] = bound
    if isinstance(p, torch.Tensor):
        return math.log(i0) - torch.log(p), m
    return np.log(i0) - np.log(p), m


def _bc_project(proj, x, m, b_hb, n_iter=100, alpha=1e-2):
    """HoughBarrier's projection onto {FP(x)[m] >= b} (teeth_recon/barrier.py:30-45, B11 kept literally)."""
    f = proj.fp(x)
    x_new = x.clone()
    for _ in range(n_iter):
        infeas = m & (f < b_hb)
        if not bool(infeas.any()):
            break
        f = torch.where(infeas, f - (b_hb + 1e-3), torch.zeros_like(f))
        x_new = x_new - alpha * proj.bp(f)
        f = proj.fp(x_new)
    return x_new


def barrier_least_squares(proj, pr, i0, bound, n_iter=200, n_biter=10, t0=0.2, t_step=0.2, alpha=0.1, beta_reg=1.0,
                          max_iter=400, do_alpha_decay=False, with_sino_barrier=True, x0_level=1e-2, on_outer=None):
    """Mono metal-artifact reduction of teeth_recon/experiments.py:651-749 on the device (defaults = its first run;
    the second run is n_iter=100, t0=0.1, max_iter=1000, do_alpha_decay=True, with_sino_barrier=False).

    pr: photon counts (A, D) or (n, A, D).  Per inner iteration: one fused FP (residual on good rays, Hough-barrier
    term on metal rays, feasibility count), one BP of their sum, one K3 step; the reference spends >= 4 FP + 2 BP
    (cost-grad FP, HoughBarrier.eval FP, bc_project FP, stats FP; two separate BPs).  The order of the reference
    is kept: step, then HoughBarrier projection, then clip to x >= 0 (src/barrier_method.py:201-206).
    Returns x with the leading shape of pr, or (x0, None)-style None if the start cannot be made feasible."""
    dev = proj.device
    p = torch.as_tensor(np.asarray(pr) if not isinstance(pr, torch.Tensor) else pr).to(dev, dtype=torch.float32).clone()
    lead = tuple(p.shape[:-2])
    p = p.reshape((-1,) + proj.sino_shape)
    n = p.shape[0]
    r, m = ray_transform_from_projection(p, bound, i0)
    bound_log = math.log(i0) - math.log(bound)
    b = torch.where(m, torch.full_like(r, bound_log), r).contiguous()
    b_hb = 0.95 * bound_log
    inv_ngood = 1.0 / (m[0].numel() - m.reshape(n, -1).sum(1)).to(torch.float32)
    m8 = m.to(torch.uint8).contiguous()
    if not with_sino_barrier:
        m_hb = torch.zeros_like(m)
    else:
        m_hb = m
    x = torch.full((n,) + proj.vol_shape, x0_level, dtype=torch.float32, device=dev)

    def fused_fp(t):
        bt = beta_reg * t if with_sino_barrier else 0.0
        u, _, st = proj.mar_fp(x, b, m8, inv_ngood, bt, b_hb)
        return u, st

    # feasibility of the start (src/barrier_method.py:128-152)
    u, st = fused_fp(t0)
    if with_sino_barrier and bool((st[:, 2] > 0).any()):
        x = _bc_project(proj, x, m_hb, b_hb)
        x.clamp_(min=0)
        u, st = fused_fp(t0)
        if bool((st[:, 2] > 0).any()):
            return None
    t = float(t0)
    for _ in range(n_biter):
        u, st = fused_fp(t)
        for _ in range(n_iter):
            g = proj.bp(u)
            if not with_sino_barrier:
                # no sinogram barrier, no feasibility test between the step and the clip: K3 does both in one pass
                # and the loop never waits for the host (the reference's second run, 8 600 of its 10 600 iterations)
                proj.barrier_update(x, g, alpha, beta_reg, t, lower_only=True, mode=_capi.UPD_STEP_CLIP)
                u, st = fused_fp(t)
                continue
            # the step's own barrier penalty -sum(log x') is +inf or nan exactly when some pixel left x > 0: the
            # reference's `clip to x >= 0` test (src/barrier_method.py:201-206) without a pass over the volume
            pen = proj.barrier_update(x, g, alpha, beta_reg, t, lower_only=True, mode=_capi.UPD_STEP_ONLY,
                                      want_penalty=True)
            u, st = fused_fp(t)                     # P at the stepped point: feasibility + next gradient
            infeasible, left_box = torch.stack([(st[:, 2] > 0).any(),
                                                ((pen == float("inf")) | torch.isnan(pen)).any()]).tolist()   # one sync
            dirty = False
            if with_sino_barrier and infeasible:
                x = _bc_project(proj, x, m_hb, b_hb)
                dirty = True
                left_box = bool((x < 0).any())
            if left_box:
                proj.barrier_update(x, None, 0.0, beta_reg, 0.0, lower_only=True, mode=_capi.UPD_CLIP_ONLY)
                dirty = True
            if dirty:
                u, st = fused_fp(t)
        if on_outer is not None:
            on_outer(dict(t=t, alpha=alpha, n_iter=n_iter, x=x, stats=st))
        t *= t_step
        if do_alpha_decay:
            alpha = alpha * math.sqrt(t_step)
            n_iter = min(int(n_iter / t_step), max_iter)
    return x.reshape(lead + proj.vol_shape)


def _checkpoint_key(proj, spectrum, meas_host, x0_host, params):
    """what a stored batch result depends on besides its own slices' data (those are fingerprinted per batch)"""
    import hashlib
    import json
    h = hashlib.sha256()
    # solver parameters that shape the result; callables (stat_cb) carry no data and their repr an address
    data_params = {k: repr(v) for k, v in sorted(params.items()) if not callable(v)}
    h.update(json.dumps({"rows": proj.rows, "cols": proj.cols, "ndet": proj.ndet, "det_width": proj.det_width,
                         "positions": getattr(proj, "positions", None), "K": spectrum.K, "E": spectrum.E,
                         "pixel_size": repr(getattr(spectrum, "pixel_size", None)), "params": data_params,
                         "meas": list(meas_host.shape), "x0": list(x0_host.shape)}, sort_keys=True).encode())
    h.update(np.ascontiguousarray(proj.angles).tobytes())
    for tab in ("source", "widths", "absorptions"):          # the assumed spectrum: another one is another job
        h.update(np.ascontiguousarray(getattr(spectrum, tab, np.zeros(0))).tobytes())
    return h.hexdigest()[:16]


def white_barrier_reconstruct(proj, spectrum, meas_host, x0_host, batch=8, rank=0, world=1, checkpoint=None, stats=None,
                              **params):
    """The C4 job: a stack of independent slices (host memory) reconstructed in resident batches on this rank's GPU.

    meas_host (S, A, D) and x0_host (S, K, N, N) are CPU tensors (pinned memory makes the copies asynchronous).  The
    rank takes the contiguous slice range ``shard_range(S, rank, world)`` -- no collective is involved -- uploads
    ``batch`` slices at a time on a copy stream while the previous batch iterates, runs ALL barrier iterations of a
    batch while it is resident (slices are independent, so nothing else ever needs to be on the device), and writes
    the result into the returned CPU tensor.  ``params`` are those of ``white_barrier_stack``.
    ``checkpoint``: a directory.  Every finished batch is written there (one .npy per batch, named after the job key --
    geometry, angles, spectrum size, solver parameters -- and the slice range, with a checksum of the batch's
    measurements); a restarted call loads the batches it finds instead of recomputing them, so a stack job that
    dies after hours resumes at the batch it was working on.  Slices are independent, so this is the whole state.
    ``stats``: a dict that receives ``slice_iters`` (int64 tensor, one entry per slice of the rank's range: the projector
    iterations actually spent on it -- early-stopped slices spend fewer than the schedule's maximum; -1 for slices
    restored from a checkpoint) and ``batches``.
    Returns (x_host for the rank's range, (start, stop))."""
    from .distributed import shard_range
    s0, s1 = shard_range(meas_host.shape[0], rank, world)
    out = torch.empty((s1 - s0,) + tuple(x0_host.shape[1:]), dtype=torch.float32,
                      pin_memory=meas_host.is_pinned())
    dev = proj.device
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    # Batches of fewer than 3 images would take the mirror-paired single-image path, whose results differ from the
    # batched kernels at rounding level; keep every batch at >= 3 images where the range allows it, so that a
    # slice's result does not depend on the batch size or on how the stack is sharded over ranks.
    K = x0_host.shape[1]
    min_slices = -(-3 // K)
    cuts = list(range(s0, s1, batch)) + [s1]
    if len(cuts) > 2 and cuts[-1] - cuts[-2] < min_slices:
        cuts[-2] = max(cuts[-3] + min(min_slices, batch), cuts[-1] - min_slices) if batch > min_slices else cuts[-2]
        if cuts[-1] - cuts[-2] < min_slices or cuts[-2] - cuts[-3] < min_slices:
            del cuts[-2]                                     # merge the short tail into the previous batch
    nxt_of = dict(zip(cuts[:-1], cuts[1:]))

    def upload(b0):
        b1 = nxt_of[b0]
        with torch.cuda.stream(copy_stream):
            m = meas_host[b0:b1].to(dev, non_blocking=True)
            x = x0_host[b0:b1].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        return b0, b1, m, x, ev

    import os
    key = _checkpoint_key(proj, spectrum, meas_host, x0_host, params) if checkpoint else None
    if checkpoint:
        os.makedirs(checkpoint, exist_ok=True)

    def ckpt_path(b0):
        return os.path.join(checkpoint, "x_%s_%06d_%06d.npy" % (key, b0, nxt_of[b0]))

    def fingerprint(b0):
        # the batch's inputs: a cheap checksum (float64 sums) guards against reusing results of other data
        m, x = meas_host[b0:nxt_of[b0]], x0_host[b0:nxt_of[b0]]
        return np.array([float(m.double().sum()), float(x.double().sum())])

    def restore(b0):
        """finished batch from an earlier run, or None"""
        if not checkpoint or not os.path.exists(ckpt_path(b0)):
            return None
        try:
            with open(ckpt_path(b0), "rb") as f:
                fp, arr = np.load(f), np.load(f)
        except Exception:
            return None                                   # truncated file of a run that died while writing
        want = (nxt_of[b0] - b0,) + tuple(x0_host.shape[1:])
        if arr.shape != want or not np.array_equal(fp, fingerprint(b0)):
            return None
        return torch.from_numpy(arr)

    def store(b0, res_host):
        tmp = ckpt_path(b0) + ".tmp"
        with open(tmp, "wb") as f:
            np.save(f, fingerprint(b0))
            np.save(f, res_host.numpy())
        os.replace(tmp, ckpt_path(b0))                    # atomic: a batch file is complete or absent

    spent = torch.full((s1 - s0,), -1, dtype=torch.int64)
    spent_dev = []
    todo = []
    for b0 in cuts[:-1]:
        done = restore(b0)
        if done is not None:
            out[b0 - s0:nxt_of[b0] - s0].copy_(done)
        else:
            todo.append(b0)
    pending = None                                        # (b0, event) of the batch whose download is in flight
    nxt = upload(todo[0]) if todo else None
    for i, b0 in enumerate(todo):
        _, b1, m, x, ev = nxt
        nxt = upload(todo[i + 1]) if i + 1 < len(todo) else None      # overlaps the iterations below
        main.wait_event(ev)
        m.record_stream(main)
        x.record_stream(main)
        res, info = white_barrier_stack(proj, spectrum, m, x, **params)
        out[b0 - s0:b1 - s0].copy_(res, non_blocking=True)
        if info is not None:
            spent_dev.append((b0, b1, info["slice_iters"]))
        if checkpoint:
            if pending is not None:                       # the previous batch's download has long finished
                pending[1].synchronize()
                store(pending[0], out[pending[0] - s0:nxt_of[pending[0]] - s0])
            done_ev = torch.cuda.Event()
            done_ev.record(main)
            pending = (b0, done_ev)
    torch.cuda.synchronize(dev)
    if checkpoint and pending is not None:
        store(pending[0], out[pending[0] - s0:nxt_of[pending[0]] - s0])
    if stats is not None:
        for b0, b1, it in spent_dev:
            spent[b0 - s0:b1 - s0] = it.cpu()
        stats["slice_iters"] = spent
        stats["batches"] = len(cuts) - 1
    return out, (s0, s1)


def mask_linear_least_squares(proj, pr, i0, bound, n_iters=200, x0=None):
    """Masked least squares of teeth_recon/experiments.py:366-401:  min_x || M (W x - r) ||^2, M = good rays.

    The reference hands (cost, grad = 2 W^T M (W x - r)) to scipy.optimize.fmin_cg (non-linear CG with a line
    search: 2-3 FP per iteration); that call keeps working unchanged over the drop-in gpu_fp / gpu_bp.  The
    objective is quadratic, so this device version runs linear CG on the normal equations W^T M W x = W^T M r
    (CGLS): exactly one FP and one BP per iteration, dot products on the device.  Both minimise the same
    functional; iterates differ, minimisers agree.  Returns (x, cost per iteration)."""
    dev = proj.device
    p = torch.as_tensor(np.asarray(pr) if not isinstance(pr, torch.Tensor) else pr).to(dev, dtype=torch.float32).clone()
    r, m = ray_transform_from_projection(p.reshape(proj.sino_shape), bound, i0)
    good = (~m).to(torch.float32)
    x = torch.zeros(proj.vol_shape, dtype=torch.float32, device=dev) if x0 is None else x0.to(dev).float().clone()
    res = good * (r - proj.fp(x))                  # M (r - W x)
    s = proj.bp(res)                               # W^T M (r - W x) = -grad/2
    d = s.clone()
    gamma = torch.sum(s.double() * s.double())
    costs = []
    for _ in range(n_iters):
        costs.append(torch.sum(res.double() ** 2).item())
        if gamma.item() == 0.0:
            break
        q = good * proj.fp(d)
        step = gamma / torch.sum(q.double() * q.double())
        x += step.float() * d
        res -= step.float() * q
        s = proj.bp(res)
        gamma_new = torch.sum(s.double() * s.double())
        d = s + (gamma_new / gamma).float() * d
        gamma = gamma_new
    return x, costs


def ineq_linear_least_squares(proj, pr, i0, bound, alpha, n_iters=200, x0=None, newton_steps=6):
    """Soft-inequality least squares of teeth_recon/experiments.py:513-582 as a device loop.

    Good rays pull W x towards the measured ray sums b; photon-starved (metal) rays, whose true ray sum is only known to
    exceed b_m = log(i0 / bound), are penalised one-sidedly: d = W x - b, d[m & (d > 0)] = 0, d[m] *= alpha.  The
    reference hands ``cost`` and ``grad = 2 W^T(d_good / n_good', alpha d_bad / n_bad)`` to scipy.optimize.fmin_cg (that
    path keeps working unchanged over the drop-in gpu_fp / gpu_bp, one H2D + D2H per projector call).  Its ``grad`` is the
    gradient of   G(x) = sum_r w_r rho_r(W x - b)_r^2,   w = 1 on good rays (1 / n_good on the two rays the reference's
    literal ``d[1 - m]`` index hits), alpha / n_bad on starved rays, rho one-sided there -- which is what fmin_cg drives
    to a stationary point.  This loop minimises G with Polak-Ribiere+ non-linear CG: G is piecewise quadratic along a
    line and W d is known after ONE forward projection of the direction, so the line search is a few Newton steps on
    sinogram-sized arrays without touching the projector -- one FP and one BP per iteration, every dot product on the
    device, no host synchronisation inside the loop.  Iterates differ from scipy's; stationary points agree.
    Returns (x, G per iteration as a float64 tensor)."""
    dev = proj.device
    p = torch.as_tensor(np.asarray(pr) if not isinstance(pr, torch.Tensor) else pr).to(dev, dtype=torch.float32).clone()
    r, m = ray_transform_from_projection(p.reshape(proj.sino_shape), bound, i0)
    b = torch.where(m, torch.full_like(r, math.log(i0) - math.log(bound)), r)
    n_bad = int(m.sum().item())
    n_good = m.numel() - n_bad
    w = torch.where(m, torch.full_like(r, float(alpha) / max(n_bad, 1)), torch.ones_like(r))
    w.view(-1)[:2] /= float(n_good)                  # ``d[1 - m] /= n_good`` indexes elements 0 and 1 (m is a 0/1 array)

    def resid(px):
        d = px - b
        return torch.where(m & (d > 0), torch.zeros_like(d), d)

    x = torch.zeros(proj.vol_shape, dtype=torch.float32, device=dev) if x0 is None else x0.to(dev).float().clone()
    px = proj.fp(x)
    d = resid(px)
    g = 2.0 * proj.bp(w * d)
    dirn = -g
    gg = torch.sum(g.double() * g.double())
    hist = []
    for _ in range(n_iters):
        hist.append(torch.sum(w.double() * d.double() ** 2))
        pd = proj.fp(dirn)
        step = torch.zeros((), dtype=torch.float64, device=dev)
        for _ in range(newton_steps):                # phi(s) = G(x + s dirn) is convex and piecewise quadratic
            cur = px + step.float() * pd - b
            act = ~(m & (cur > 0))
            dd = torch.where(act, cur, torch.zeros_like(cur))
            g1 = torch.sum((w * dd).double() * pd.double())
            g2 = torch.sum((w * act).double() * pd.double() ** 2)
            step = step - torch.where(g2 > 0, g1 / g2, torch.zeros_like(g1))
        x += step.float() * dirn
        px += step.float() * pd
        d = resid(px)
        g_new = 2.0 * proj.bp(w * d)
        gg_new = torch.sum(g_new.double() * g_new.double())
        beta = torch.clamp((gg_new - torch.sum(g_new.double() * g.double())) / gg, min=0.0)     # Polak-Ribiere+
        beta = torch.where(torch.isfinite(beta), beta, torch.zeros_like(beta))
        dirn = -g_new + beta.float() * dirn
        g, gg = g_new, gg_new
    hist.append(torch.sum(w.double() * d.double() ** 2))
    return x, torch.stack(hist)
