"""Measurement synthesis on the device, so that C3/C4 inputs never round-trip through the host.

  create_projection_with_poisson_noise   teeth_recon/experiments.py:196-200
  white_sinogram                         WhiteProjection.calc_sinogram_with_gt, src/white_projection.py:100-102
"""
import math

import torch


def create_projection_with_poisson_noise(i0, proj, v, generator=None):
    """p ~ Poisson(i0 * exp(-W v)) as float32 (the reference draws with numpy's global RNG; here a torch generator
    on the device -- the distribution is the same, the stream of random numbers is not)."""
    r = proj.fp(v)
    lam = torch.exp(math.log(i0) - r.double())
    return torch.poisson(lam, generator=generator).to(torch.float32)


def white_sinogram(proj, spectrum, gt):
    """I(gt): the synthetic measurement of the white-beam model, one fused K1 launch per group of images."""
    inten, _, _ = proj.white_fp(spectrum, gt, want_qmu=False, want_loss=False)
    return inten
