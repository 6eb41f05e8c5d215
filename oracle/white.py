"""Oracle restatement of the white-beam (polychromatic) model -- TEST INFRASTRUCTURE.

Follows ``src/white_projection.py`` of the reference line by line in float64 numpy around the float32
oracle projector, exactly like the reference runs float64 numpy around float32 ASTRA:

  integrate            src/white_projection.py:11-16
  source normalisation src/white_projection.py:69-73
  FP_white             src/white_projection.py:86-98
  calc_mu              src/white_projection.py:130-139
  BP_white             src/white_projection.py:104-113
  func / grad          src/white_projection.py:116-127
  regularisers         src/white_rec.py:180-194  (calc_uneq_reg_grad, calc_triv_conc_grad)
  Iteration            src/white_rec.py:199-244

Pinned by tests/golden/white_*.npz, which were produced by executing the reference's own
``white_projection.py`` on an oracle-backed ``astra_proxy`` (tests/golden/make_golden.py).
"""
import numpy as np

from . import Geometry, fp as _fp, bp as _bp


def integrate(ar, energy_grid):
    """sum_e ar[..., e] * w_e with w = energy_grid[:, 1] (src/white_projection.py:11-16)."""
    res_shape = ar.shape[:-1]
    num_en = energy_grid.shape[0]
    return np.sum(ar.reshape(-1, num_en) * energy_grid[:, 1].reshape(1, -1), axis=1).reshape(res_shape)


class WhiteOracle(object):
    def __init__(self, source, pixel_size, element_absorptions, energy_grid, ph_size, n_angles=360,
                 geometry=None):
        self.pixel_size = pixel_size
        self.element_absorptions = np.asarray(element_absorptions, dtype=np.float64)
        self.energy_grid = np.asarray(energy_grid, dtype=np.float64)
        self.geom = geometry if geometry is not None else Geometry(
            ph_size[0], ph_size[1], int(np.sqrt(2) * ph_size[0] + 1), np.linspace(0, np.pi, n_angles))
        self.proj_shape = self.geom.sino_shape
        self.conc_shape = (self.element_absorptions.shape[0],) + self.geom.vol_shape
        source = np.asarray(source, dtype=np.float64)
        self.source = source / integrate(source, self.energy_grid)      # :69-71
        self.sinogram = None

    def FP(self, x):
        return _fp(self.geom, x)

    def BP(self, x):
        return _bp(self.geom, x)

    def FP_white(self, c):                                               # :86-98
        K = len(c)
        conc_fp = np.array([self.FP(cc) for cc in c])
        flat_fp = conc_fp.reshape(K, -1, 1)
        ea = self.element_absorptions.reshape(K, 1, -1)
        exp_arg = -self.pixel_size * np.sum(flat_fp * ea, axis=0)
        exp = np.exp(exp_arg)
        integral = integrate(self.source.reshape(1, -1) * exp, self.energy_grid)
        return integral, exp_arg, exp, flat_fp

    def calc_sinogram_with_gt(self, gt):                                 # :100-102
        self.sinogram = self.FP_white(gt)[0]

    def calc_mu(self, exp):                                              # :130-139
        src = self.source.reshape(1, -1)
        return np.array([-integrate(src * self.pixel_size * exp * ea.reshape(1, -1), self.energy_grid)
                         for ea in self.element_absorptions])

    def BP_white(self, Q, exp):                                          # :104-113
        mu = self.calc_mu(exp)
        bp_arg = Q * mu
        conc = np.zeros(self.conc_shape, dtype=np.float64)
        for k in range(conc.shape[0]):
            conc[k] = -self.BP(bp_arg[k].reshape(self.proj_shape))
        return conc, mu

    def func(self, c):                                                   # :116-120
        fp = self.FP_white(c)[0]
        q = self.sinogram - fp
        return 0.5 * np.linalg.norm(q) ** 2

    def grad(self, c):                                                   # :122-127
        fp, _, exp, _ = self.FP_white(c)
        q = self.sinogram - fp
        return self.BP_white(q, exp)[0]


class ChunkedWhiteOracle(WhiteOracle):
    """Same arithmetic as WhiteOracle, ray by ray, with the (rays, E) float64 temporaries of FP_white / calc_mu built
    for `chunk` rays at a time (at 2048^2 x 1800 x 64 bins the reference's temporaries are 5.3 + 3 x 2.7 GB).  Every
    operation is element-wise or a sum along E of one ray, so func and grad are bit-identical to the unchunked class
    (tests/test_oracle_golden.py checks that)."""

    chunk = 1 << 18

    def _model(self, c, want_mu):
        K = len(c)
        conc_fp = np.array([self.FP(cc) for cc in c]).reshape(K, -1)
        R = conc_fp.shape[1]
        integral = np.empty(R)
        mu = np.empty((K, R)) if want_mu else None
        ea = self.element_absorptions.reshape(K, 1, -1)
        src = self.source.reshape(1, -1)
        for r0 in range(0, R, self.chunk):
            r1 = min(R, r0 + self.chunk)
            flat_fp = conc_fp[:, r0:r1].reshape(K, -1, 1)
            exp = np.exp(-self.pixel_size * np.sum(flat_fp * ea, axis=0))
            integral[r0:r1] = integrate(src * exp, self.energy_grid)
            if want_mu:
                for k, eak in enumerate(self.element_absorptions):
                    mu[k, r0:r1] = -integrate(src * self.pixel_size * exp * eak.reshape(1, -1), self.energy_grid)
        return integral, mu

    def calc_sinogram_with_gt(self, gt):
        self.sinogram = self._model(gt, False)[0]

    def func(self, c):
        q = self.sinogram - self._model(c, False)[0]
        return 0.5 * np.linalg.norm(q) ** 2

    def grad(self, c):
        fp, mu = self._model(c, True)
        bp_arg = (self.sinogram - fp) * mu
        conc = np.zeros(self.conc_shape, dtype=np.float64)
        for k in range(conc.shape[0]):
            conc[k] = -self.BP(bp_arg[k].reshape(self.proj_shape))
        return conc


def calc_uneq_reg_grad(c):
    """d|c0 c1|/dc_i = sum_{j != i} c_j * sign(c0 c1)  (src/white_rec.py:180-186)."""
    out = np.zeros_like(c)
    for i in range(len(c)):
        out[i] = (c[np.arange(len(c)) != i] * np.sign(c[0] * c[1])).sum(axis=0)
    return out


def calc_triv_conc_grad(c):
    """c*(c*(4c-3)+1)  (src/white_rec.py:188-194; the code, not its comment)."""
    res = 4 * c.copy()
    res = c * (res - 3)
    return c * (res + 1)


def iteration(wo, c, alpha, beta_reg, clip_concentrations=False, dissimilarity_constraint=False,
              triv_conc_constraint=False, iter_num=None):
    """One step of src/white_rec.py:199-244 (batch filtering off, its default at :342)."""
    fp, _, exp, _ = wo.FP_white(c)
    q = wo.sinogram - fp
    bp_grad, mu = wo.BP_white(q, exp)
    reg_grad = 0
    reg_loss = 0
    if dissimilarity_constraint:
        reg_grad = calc_uneq_reg_grad(c)
        reg_loss += np.linalg.norm(beta_reg * c[0] * c[1])
    tc_grad = 0
    if triv_conc_constraint:
        tc_grad = calc_triv_conc_grad(c)
    step = alpha * (bp_grad + beta_reg * (reg_grad + tc_grad))
    if iter_num is not None:
        if iter_num % 4 == 0:
            step[1] = 0
        else:
            step[0] = 0
    c = c - step
    full_loss = np.linalg.norm(q) + reg_loss
    if clip_concentrations:
        c = np.clip(c, 0, 1)
    return c, full_loss
