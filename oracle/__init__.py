"""CPU oracle for the whitomo hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``whitomo_b200`` never does.

Contents
--------
* ``astra_linear.c`` (-> ``liboracle.so``): restatement of the ASTRA CPU ``linear`` projector FP/BP,
  SIRT and FBP that the reference reaches through ``src/astra_proxy.py:35-102``.
* ``white.py``: numpy float64 restatement of ``src/white_projection.py:11-139`` (white-beam model).
* ``barrier.py``: py3 restatement of ``src/barrier_method.py:14-240``, ``teeth_recon/barrier.py:10-69``
  and the mono MAR closures ``teeth_recon/experiments.py:290-293,651-749``.

Parity pins: G1/G2 golden artifacts for FP/BP/FBP (tests/test_oracle_golden.py); the white-beam and
barrier restatements are pinned against outputs of the reference's own Python modules executed in the
build container on an oracle-backed ``astra`` (tests/golden/make_golden.py).  SIRT: parity unpinned.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "astra_linear.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        geom = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, f32p, ctypes.c_int]
        for name in ("orc_fp", "orc_bp", "orc_fbp"):
            getattr(L, name).argtypes = geom + [f32p, f32p]
            getattr(L, name).restype = ctypes.c_int
        for name in ("orc_fp_mt", "orc_bp_mt"):
            getattr(L, name).argtypes = geom + [f32p, f32p, ctypes.c_int]
            getattr(L, name).restype = ctypes.c_int
        for name in ("orc_fp_f64", "orc_bp_f64"):
            getattr(L, name).argtypes = geom + [f64p, f64p]
            getattr(L, name).restype = ctypes.c_int
        for name in ("orc_fp_f64_mt", "orc_bp_f64_mt"):
            getattr(L, name).argtypes = geom + [f64p, f64p, ctypes.c_int]
            getattr(L, name).restype = ctypes.c_int
        L.orc_sirt.argtypes = geom + [f32p, f32p, ctypes.c_int]
        L.orc_sirt.restype = ctypes.c_int
        L.orc_ramp_filter.argtypes = [ctypes.c_int, ctypes.c_int, f32p, f32p]
        L.orc_ramp_filter.restype = ctypes.c_int
        L.orc_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _p64(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class Geometry(object):
    """(rows, cols, ndet, det_width, angles[float32]) -- mirrors what create_proj_geom/create_vol_geom hold."""

    def __init__(self, rows, cols, ndet, angles, det_width=1.0):
        self.rows, self.cols, self.ndet = int(rows), int(cols), int(ndet)
        self.det_width = float(det_width)
        # ASTRA stores projection angles as float32
        self.angles = np.ascontiguousarray(np.asarray(angles, dtype=np.float64).astype(np.float32))
        self.nangles = int(self.angles.shape[0])

    @classmethod
    def standard(cls, n, n_angles):
        """The reference's standard set-up: src/white_projection.py:41-47,60."""
        return cls(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, n_angles))

    def _args(self):
        return (self.rows, self.cols, self.ndet, self.det_width, _p32(self.angles), self.nangles)

    @property
    def vol_shape(self):
        return (self.rows, self.cols)

    @property
    def sino_shape(self):
        return (self.nangles, self.ndet)


def fp(g, vol, threads=1):
    """p = W x, float32, ASTRA CPU 'FP' + 'linear' projector (src/astra_proxy.py:35-49)."""
    v = _f32(vol).reshape(g.vol_shape)
    out = np.empty(g.sino_shape, dtype=np.float32)
    if threads == 1:
        rc = lib().orc_fp(*g._args(), _p32(v), _p32(out))
    else:
        rc = lib().orc_fp_mt(*g._args(), _p32(v), _p32(out), int(threads))
    assert rc == 0
    return out


def bp(g, sino, threads=1):
    """g = W^T q, float32 scatter, ASTRA CPU 'BP' (src/astra_proxy.py:52-66)."""
    s = _f32(sino).reshape(g.sino_shape)
    out = np.empty(g.vol_shape, dtype=np.float32)
    if threads == 1:
        rc = lib().orc_bp(*g._args(), _p32(s), _p32(out))
    else:
        rc = lib().orc_bp_mt(*g._args(), _p32(s), _p32(out), int(threads))
    assert rc == 0
    return out


def fp64(g, vol, threads=1):
    """float64 closed-form arbiter of W x (no running sums); threads=0: all host cores, same result"""
    v = np.ascontiguousarray(vol, dtype=np.float64).reshape(g.vol_shape)
    out = np.empty(g.sino_shape, dtype=np.float64)
    if threads == 1:
        assert lib().orc_fp_f64(*g._args(), _p64(v), _p64(out)) == 0
    else:
        assert lib().orc_fp_f64_mt(*g._args(), _p64(v), _p64(out), int(threads)) == 0
    return out


def bp64(g, sino, threads=1):
    """float64 closed-form arbiter of W^T q (pixel-driven gather form); threads=0: all host cores, same result"""
    s = np.ascontiguousarray(sino, dtype=np.float64).reshape(g.sino_shape)
    out = np.empty(g.vol_shape, dtype=np.float64)
    if threads == 1:
        assert lib().orc_bp_f64(*g._args(), _p64(s), _p64(out)) == 0
    else:
        assert lib().orc_bp_f64_mt(*g._args(), _p64(s), _p64(out), int(threads)) == 0
    return out


def sirt(g, sino, n_iters=100):
    """ASTRA CPU 'SIRT' (src/astra_proxy.py:69-84).  PARITY UNPINNED (no ASTRA SIRT artifact in the reference)."""
    s = _f32(sino).reshape(g.sino_shape)
    out = np.empty(g.vol_shape, dtype=np.float32)
    assert lib().orc_sirt(*g._args(), _p32(s), _p32(out), int(n_iters)) == 0
    return out


def ramp_filter(g, sino):
    s = _f32(sino).reshape(g.sino_shape)
    out = np.empty_like(s)
    assert lib().orc_ramp_filter(g.ndet, g.nangles, _p32(s), _p32(out)) == 0
    return out


def fbp(g, sino):
    """ASTRA CPU 'FBP' (src/astra_proxy.py:87-102); pinned by src/fbp.nptxt."""
    s = _f32(sino).reshape(g.sino_shape)
    out = np.empty(g.vol_shape, dtype=np.float32)
    assert lib().orc_fbp(*g._args(), _p32(s), _p32(out)) == 0
    return out


def num_threads():
    return int(lib().orc_num_threads())
