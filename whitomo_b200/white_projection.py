"""Drop-in for ``src/white_projection.py``: the polychromatic ("white-beam") forward model and its gradient.

Same constructor, attributes and methods as the reference class; the projector work and -- for ``func`` /
``grad`` -- the whole point-wise spectral model run in the sm_100a kernels:

  I[r]   = sum_e s_e w_e exp(-pix sum_k ea[k,e] P_k[r]),  P_k = W c_k      (src/white_projection.py:86-98)
  mu_k[r] = -pix sum_e s_e w_e ea[k,e] exp(...)                              (src/white_projection.py:130-139)
  func   = 1/2 ||s - I||^2,   grad_k = -W^T((s - I) mu_k)                   (src/white_projection.py:104-127)

``func``/``grad`` never materialise the reference's (K, rays, E) / (rays, E) float64 temporaries: K1's epilogue
produces I, the residual and q*mu_k per ray, K2 backprojects them.  ``FP_white`` / ``calc_mu`` / ``BP_white`` keep
the reference's array-returning signatures (numpy float64 around float32 projections, like the reference's
numpy around ASTRA) for callers that want the intermediates.

Deviations, all behind the same signature: ``sinogram=None`` is accepted (B1); the caller's
sinogram is copied, not normalised in place (B2).
"""
from __future__ import print_function, division

import numpy as np
import torch

from .astra_proxy import *  # noqa: F401,F403  (re-exports `astra`, gpu_fp, gpu_bp like the reference)
from .astra_proxy import astra, gpu_fp, gpu_bp
from .projector import Spectrum


def integrate(ar, energy_grid):
    """sum_e ar[..., e] * w_e, w = energy_grid[:, 1]  (src/white_projection.py:11-16)."""
    ar = np.asarray(ar)
    w = np.asarray(energy_grid)[:, 1]
    return np.tensordot(ar.reshape(-1, w.shape[0]), w, axes=([1], [0])).reshape(ar.shape[:-1])


class WhiteProjection:
    def __init__(self, source, pixel_size, element_numbers, element_absorptions,
                 energy_grid, ph_size, sinogram=None, n_angles=360, angles=None):
        self.dt = np.float64
        self.source_init = np.asarray(source, dtype=np.float64)
        self.pixel_size = pixel_size
        self.element_numbers = element_numbers
        self.element_absorptions = np.asarray(element_absorptions, dtype=np.float64)
        self.ph_size = tuple(ph_size) if ph_size is not None else None
        self.n_angles = n_angles
        self.energy_grid = np.asarray(energy_grid, dtype=np.float64)
        self.angles = angles
        self.sinogram = None if sinogram is None else np.array(sinogram, dtype=np.float64).ravel()

        if sinogram is not None:
            self.ph_size = (sinogram.shape[1], sinogram.shape[1])           # :35-36

        if self.angles is None:                                              # :39-47
            det = int(np.sqrt(2) * self.ph_size[0] + 1)
            self.proj_geom = astra.create_proj_geom('parallel', 1.0, det, np.linspace(0, np.pi, n_angles))
        else:                                                                # :48-58 (real-data branch)
            w = int((self.ph_size[0] - 1) / np.sqrt(2))
            self.n_angles = np.asarray(angles).shape[0]
            self.proj_geom = astra.create_proj_geom('parallel', 1.0, self.ph_size[0], self.angles)
            self.ph_size = (w, w)

        self.vol_geom = astra.create_vol_geom(*self.ph_size)
        self.projector = astra.create_projector('linear', self.proj_geom, self.vol_geom)
        self.proj_shape = (self.n_angles, self.proj_geom['DetectorCount'])
        self.conc_shape = (len(self.element_numbers), self.ph_size[0], self.ph_size[1])

        sum_energy = self.integrate(self.source_init)                        # :69-73
        self.source = self.source_init / sum_energy
        if self.sinogram is not None:
            self.sinogram = self.sinogram / sum_energy

        # device side: geometry handle, spectrum tables, cached measurement
        self._proj = astra.projector_for(self.proj_geom, self.vol_geom, self.projector)
        self._spectrum = Spectrum(self.source, self.energy_grid[:, 1], self.element_absorptions, self.pixel_size)
        self._meas_dev = None
        self._meas_src = None

    # ---- reference API ------------------------------------------------------------------------
    def integrate(self, ar):
        return integrate(ar, self.energy_grid)

    def FP(self, x):
        return gpu_fp(self.proj_geom, self.vol_geom, x, self.projector)

    def BP(self, x):
        return gpu_bp(self.proj_geom, self.vol_geom, x, self.projector)

    def FP_white(self, c):
        """(Integral, exp_arg, exp, flat_fp) with the reference's shapes and dtypes (:86-98)."""
        K = len(c)
        c_dev = self._to_dev(c)
        flat_fp = self._proj.fp(c_dev).cpu().numpy().astype(np.float64).reshape(K, -1, 1)
        exp_arg = -self.pixel_size * np.einsum('kr,ke->re', flat_fp[:, :, 0], self.element_absorptions)
        exp = np.exp(exp_arg)
        return self.integrate(self.source.reshape(1, -1) * exp), exp_arg, exp, flat_fp

    def calc_sinogram_with_gt(self, gt):
        self.gt = gt
        inten, _, _ = self._proj.white_fp(self._spectrum, self._to_dev(gt), want_qmu=False, want_loss=False)
        self._meas_dev = inten[0].contiguous()
        self.sinogram = self._meas_dev.cpu().numpy().astype(np.float64).ravel()
        self._meas_src = self.sinogram

    def calc_mu(self, exp):
        s = (self.source * self.energy_grid[:, 1]).reshape(1, -1)
        return np.array([-(self.pixel_size * exp * s * ea.reshape(1, -1)).sum(axis=1)
                         for ea in self.element_absorptions])

    def BP_white(self, Q, exp):
        mu = self.calc_mu(exp)
        bp_arg = (np.asarray(Q).reshape(1, -1) * mu).reshape((-1,) + self.proj_shape)
        conc = self._proj.bp(torch.from_numpy(bp_arg.astype(np.float32)), scale=-1.0).cpu().numpy().astype(self.dt)
        return conc.reshape(self.conc_shape), mu

    def func(self, concentrations):
        """1/2 ||s - I(c)||^2 -- one fused K1 launch per element pair, scalar read back (:116-120)."""
        _, _, loss = self._proj.white_fp(self._spectrum, self._to_dev(concentrations), meas=self._meas(),
                                         want_intensity=False, want_qmu=False)
        return float(loss[0].item())

    def grad(self, concentrations):
        """-W^T((s - I) mu_k) -- K1 (fused epilogue) then K2 (:122-127).  Returns what it was given: numpy in ->
        numpy float64 out, CUDA tensor in -> CUDA float32 tensor out."""
        g = self.grad_dev(self._to_dev(concentrations))
        if isinstance(concentrations, torch.Tensor):
            return g
        return g.cpu().numpy().astype(self.dt)

    # ---- device-resident helpers used by the solvers --------------------------------------------
    def func_grad_dev(self, c_dev):
        """(loss (S,) float64 tensor, grad like c_dev) without leaving the device."""
        _, qmu, loss = self._proj.white_fp(self._spectrum, c_dev, meas=self._meas(), want_intensity=False)
        g = self._proj.bp(qmu.reshape((-1,) + self.proj_shape), scale=-1.0)
        return loss, g.reshape(c_dev.shape)

    def grad_dev(self, c_dev):
        return self.func_grad_dev(c_dev)[1]

    def _to_dev(self, c):
        if isinstance(c, torch.Tensor):
            return c.to(self._proj.device, dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(c, dtype=np.float32)).to(self._proj.device)

    def _meas(self):
        if self.sinogram is None:
            raise ValueError("no sinogram: pass one to the constructor or call calc_sinogram_with_gt")
        if self._meas_dev is None or self._meas_src is not self.sinogram:
            self._meas_dev = torch.from_numpy(np.ascontiguousarray(self.sinogram, dtype=np.float32)
                                              .reshape(self.proj_shape)).to(self._proj.device)
            self._meas_src = self.sinogram
        return self._meas_dev
