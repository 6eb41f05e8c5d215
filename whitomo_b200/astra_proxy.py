"""Drop-in for the reference's operator boundary ``src/astra_proxy.py`` (same names, argument meaning and
return types), running on the hand-written sm_100a kernels instead of ASTRA's CPU algorithms.

  gpu_fp   src/astra_proxy.py:35-49     numpy in -> numpy (A, D) float32 out
  gpu_bp   src/astra_proxy.py:52-66     ``supersampling`` accepted and ignored, like the reference (B16)
  cpu_sirt src/astra_proxy.py:69-84
  cpu_fbp  src/astra_proxy.py:87-102    ``n_iters`` accepted and ignored (B16)
  AstraProxy  src/astra_proxy.py:9-32   the 2D/3D name table (only dims == 2 is ever exercised)
  astra    re-exported, because ``from astra_proxy import *`` is how white_projection.py gets it

The numpy boundary costs one H2D and one D2H copy per call, exactly where the reference copies into and out
of ASTRA's object store; solvers that stay on the device use ``whitomo_b200.projector.Projector`` directly.
If a torch CUDA tensor is passed instead of a numpy array, a CUDA tensor is returned and nothing is copied.
"""
import numpy as np
import torch

from . import astra_shim as astra

__all__ = ["astra", "AstraProxy", "gpu_fp", "gpu_bp", "cpu_sirt", "cpu_fbp"]


_ALGO_ATTRS = ("fp", "bp", "sirt", "fbp")


class AstraProxy:
    """Name table keyed by dimensionality; attribute names as in src/astra_proxy.py:9-32."""

    def __init__(self, dims):
        if dims != 2:
            # dims == 3 maps to ASTRA's *3D algorithms in the reference, which it never runs
            raise NotImplementedError("only 2D data is on the hot path (dims=%r)" % (dims,))
        store = astra.data2d
        self.data_create, self.data_delete, self.data_get = store.create, store.delete, store.get
        for a in _ALGO_ATTRS:
            setattr(self, a + "_algo_name", a.upper())
        self.dart_mask_name, self.dart_smoothing_name = "DARTMASK", "DARTSMOOTHING"


def _out(t, like):
    if isinstance(like, torch.Tensor):
        return t
    return t.cpu().numpy()


def _check(arr, shape, what):
    if tuple(arr.shape) != tuple(shape):
        raise ValueError("The dimensions of the %s do not match those specified in the geometry: %s != %s"
                         % (what, tuple(arr.shape), tuple(shape)))


def gpu_fp(pg, vg, v, proj_id):
    AstraProxy(len(v.shape))
    p = astra.projector_for(pg, vg, proj_id)
    _check(v, p.vol_shape, "volume")
    return _out(p.fp(v), v)


def gpu_bp(pg, vg, rt, proj_id, supersampling=1):
    AstraProxy(len(rt.shape))
    p = astra.projector_for(pg, vg, proj_id)
    _check(rt, p.sino_shape, "sinogram")
    return _out(p.bp(rt), rt)


def cpu_sirt(pg, vg, proj_id, sm, n_iters=100):
    AstraProxy(len(sm.shape))
    p = astra.projector_for(pg, vg, proj_id)
    _check(sm, p.sino_shape, "sinogram")
    return _out(p.sirt(sm, n_iters), sm)


def cpu_fbp(pg, vg, proj_id, sm, n_iters=100):
    AstraProxy(len(sm.shape))
    p = astra.projector_for(pg, vg, proj_id)
    _check(sm, p.sino_shape, "sinogram")
    return _out(p.fbp(sm), sm)
