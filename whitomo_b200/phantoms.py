"""Synthetic inputs for the hot path.  The reference's own inputs need xraylib / h5 data that are not
available; the recipes and pinned constants they were made from are reproduced here at any size:

  two_discs        src/button_generator.py:24-41 (discs at 0.3125N / 0.78125N, radii 0.195N / 0.078N)
  synth_spectrum   src/phantoms.py:159-181 (2-bin set) and its E-bin generalisation after src/synth_spec.py:183-196
  tooth_phantom    teeth_recon/experiments.py:35-86 with the levels pinned at misc/calc_recon_areas.py:15
"""
import numpy as np

CA_LEVEL, AU_LEVEL = 0.10358164, 9.21754054      # mu*pixel of Ca / Au at 45 keV, misc/calc_recon_areas.py:15


def two_discs(n, z=None):
    """(2, n, n) binary element maps; `z` in [0,1) scales the radii by 0.8 + 0.4 z so that slices differ (None: reference size)."""
    yy, xx = np.mgrid[0:n, 0:n]
    s = 1.0 if z is None else 0.8 + 0.4 * z
    c = np.zeros((2, n, n), dtype=np.float32)
    c[0][(xx - 0.3125 * n) ** 2 + (yy - 0.3125 * n) ** 2 <= (0.195 * n * s) ** 2] = 1.0
    c[1][(xx - 0.78125 * n) ** 2 + (yy - 0.78125 * n) ** 2 <= (0.078 * n * s) ** 2] = 1.0
    return c


def synth_spectrum(n_bins=64, n=256):
    """(energy_grid (E,2), source (E,), element_absorptions (2,E), pixel_size).  n_bins == 2 is exactly the
    reference's get_synth_data_small; otherwise energies linspace(0.5, 16.25, E), two Gaussian lines 2:1 at
    2.5 / 10 keV (sigma 0.5) + 5% continuum, log-linear absorption curves through the pinned peak values
    4000@2.5 / 302.04@10 (element 0) and 300@2.5 / 2000@10 (element 1); pixel_size keeps the optical depth."""
    if n_bins == 2:
        grid = np.array([[2.5, 10], [0.005, 0.005]]).T
        return grid, np.array([2.0, 1.0]), np.stack([np.array([4000, 302.04159662]), np.array([300, 2000])]), 1e-6 * 256 / n
    e = np.linspace(0.5, 16.25, n_bins)
    w = np.full(n_bins, e[1] - e[0])
    source = 2.0 * np.exp(-0.5 * ((e - 2.5) / 0.5) ** 2) + 1.0 * np.exp(-0.5 * ((e - 10.0) / 0.5) ** 2) + 0.05

    def curve(p25, p10):
        # log-linear through the two pinned points
        slope = (np.log(p10) - np.log(p25)) / (10.0 - 2.5)
        return np.exp(np.log(p25) + slope * (e - 2.5))

    ea = np.stack([curve(4000.0, 302.04159662), curve(300.0, 2000.0)])
    return np.stack([e, w]).T, source, ea, 1e-6 * 256 / n


def tooth_phantom(n, implant=True):
    """teeth_recon/experiments.py:35-86 (three 'teeth' + crescent implant), levels in units of mu*pixel."""
    phantom = np.zeros((n, n))
    half, sq = n / 2, n / 14
    y, x = np.meshgrid(range(n), range(n))
    xx = (x - half).astype('float32')
    yy = (y - half).astype('float32')
    r = np.sqrt(xx * xx + yy * yy)
    th = np.arctan2(yy, xx)
    tooth = r <= sq * (1 + np.cos(2 * th) + np.sin(2 * th) ** 2)
    tooth = tooth | ((xx * xx + yy * yy) <= (0.09 * n) ** 2)
    tooth = tooth | np.roll(tooth, n // 3, axis=0) | np.roll(tooth, -n // 3, axis=0)
    phantom[tooth] = CA_LEVEL
    imp = (xx / (0.11 * n)) ** 2 + (yy / (0.07 * n)) ** 2 < 1
    imp &= y <= half
    imp &= ((xx / (0.11 * n)) ** 2 + ((yy - 0.025 * n) / (0.07 * n)) ** 2) > 1
    if implant:
        phantom[imp] = AU_LEVEL
    return phantom.astype(np.float32)
