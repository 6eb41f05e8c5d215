"""The reusable parts of ``src/white_rec.py`` without its import-time side effects:
the two regulariser gradients and one step of the plain gradient-descent ``Iteration``.
Functions accept numpy arrays or torch tensors (the arithmetic is element-wise host/torch glue; the
projector work inside ``iteration`` runs in the sm_100a kernels through WhiteProjection)."""
import numpy as np
import torch


def calc_uneq_reg_grad(c):
    """d|c0 c1|/dc_i = c_{j != i} * sign(c0 c1)   (src/white_rec.py:180-186; K == 2 like the reference's use)."""
    xp = torch if isinstance(c, torch.Tensor) else np
    sgn = xp.sign(c[0] * c[1])
    out = []
    for i in range(len(c)):
        others = [c[j] * sgn for j in range(len(c)) if j != i]
        acc = others[0]
        for o in others[1:]:
            acc = acc + o
        out.append(acc)
    return xp.stack(out)


def calc_triv_conc_grad(c):
    """c*(c*(4c-3)+1)  (src/white_rec.py:188-194 -- the code, not the polynomial in its comment; B13)."""
    return c * (c * (4 * c - 3) + 1)


def iteration(wp, c, alpha, beta_reg, batch_filtering=False, clip_concentrations=False,
              dissimilarity_constraint=False, triv_conc_constraint=False, iter_num=None):
    """One step of ``Iteration`` (src/white_rec.py:199-244) on the device.  ``wp`` is a WhiteProjection with a
    sinogram; ``c`` a (K, N, N) CUDA tensor.  Returns (c_new, full_loss) with
    full_loss = ||q|| + ||beta c0 c1||  (:236-237).  Batch masks (:225-227) are off by default in the reference
    (:342) and not implemented."""
    if batch_filtering:
        raise NotImplementedError("batch filtering is disabled in the reference's main loop and out of scope")
    loss, bp_grad = wp.func_grad_dev(c)
    reg_grad = 0
    reg_loss = 0.0
    if dissimilarity_constraint:
        reg_grad = calc_uneq_reg_grad(c)
        reg_loss = torch.linalg.vector_norm(beta_reg * c[0] * c[1]).item()
    tc_grad = calc_triv_conc_grad(c) if triv_conc_constraint else 0
    step = alpha * (bp_grad + beta_reg * (reg_grad + tc_grad))
    if iter_num is not None:
        if iter_num % 4 == 0:
            step[1] = 0
        else:
            step[0] = 0
    c = c - step
    full_loss = float(np.sqrt(2.0 * loss[0].item())) + reg_loss
    if clip_concentrations:
        c = torch.clamp(c, 0, 1)
    return c, full_loss
