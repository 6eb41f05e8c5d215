"""A CPU stand-in with the Projector interface, backed by the oracle -- lets the host-side decomposition logic
(whitomo_b200.distributed) run under gloo without a GPU.  Test infrastructure only."""
import numpy as np
import torch

import oracle
from oracle import white as ow


class OracleProjector(object):
    def __init__(self, rows, cols, ndet, angles):
        self.geom = oracle.Geometry(rows, cols, ndet, angles)
        self.rows, self.cols, self.ndet, self.nangles = rows, cols, ndet, self.geom.nangles
        self.vol_shape, self.sino_shape = self.geom.vol_shape, self.geom.sino_shape
        self.device = torch.device("cpu")

    def _sub(self, rng):
        a0, a1 = (0, self.nangles) if rng is None else rng
        return a0, a1, oracle.Geometry(self.rows, self.cols, self.ndet, self.geom.angles[a0:a1].astype(np.float64))

    def fp(self, vol, angle_range=None):
        v = vol.numpy().reshape((-1,) + self.vol_shape)
        a0, a1, g = self._sub(angle_range)
        out = np.zeros((v.shape[0],) + self.sino_shape, np.float32)
        if a1 > a0:
            for i in range(v.shape[0]):
                out[i, a0:a1] = oracle.fp(g, v[i])
        return torch.from_numpy(out.reshape(tuple(vol.shape[:-2]) + self.sino_shape))

    def bp(self, sino, angle_range=None, scale=1.0):
        s = sino.numpy().reshape((-1,) + self.sino_shape)
        a0, a1, g = self._sub(angle_range)
        out = np.zeros((s.shape[0],) + self.vol_shape, np.float32)
        if a1 > a0:
            for i in range(s.shape[0]):
                out[i] = scale * oracle.bp(g, s[i, a0:a1])
        return torch.from_numpy(out.reshape(tuple(sino.shape[:-2]) + self.vol_shape))

    def bp_row_blocks(self, nimages, nblocks):
        """same contract as Projector.bp_row_blocks (no mirror pairing, tile height 4)"""
        per = -(-self.rows // max(1, nblocks))
        per = -(-per // 4) * 4
        return [(b0, min(self.rows, b0 + per), [(b0, min(self.rows, b0 + per))]) for b0 in range(0, self.rows, per)]

    def bp_rows(self, sino, out, row_begin, row_end, angle_range=None, scale=1.0, reuse_packed=False):
        full = self.bp(sino, angle_range=angle_range, scale=scale).reshape(out.shape)
        out[:, row_begin:row_end] = full[:, row_begin:row_end]
        return out

    def white_fp(self, spectrum, conc, meas=None, want_intensity=True, want_qmu=True, want_loss=True, angle_range=None):
        """spectrum: an oracle.white.WhiteOracle carrying the tables"""
        c = conc.numpy().astype(np.float64)
        a0, a1, g = self._sub(angle_range)
        K = c.shape[0]
        inten = np.zeros(self.sino_shape, np.float32)
        qmu = np.zeros((K,) + self.sino_shape, np.float32)
        loss = 0.0
        if a1 > a0:
            wo = ow.WhiteOracle(spectrum.source, spectrum.pixel_size, spectrum.element_absorptions,
                                spectrum.energy_grid, self.vol_shape, geometry=g)
            wo.source = spectrum.source
            I, _, ex, _ = wo.FP_white(c)
            q = (meas.numpy()[a0:a1].ravel() if meas is not None else 0.0) - I
            inten[a0:a1] = I.reshape(a1 - a0, -1)
            qmu[:, a0:a1] = (q * wo.calc_mu(ex)).reshape(K, a1 - a0, -1)
            loss = 0.5 * np.sum(q ** 2)
        return (torch.from_numpy(inten)[None], torch.from_numpy(qmu)[None],
                torch.tensor([loss], dtype=torch.float64))
