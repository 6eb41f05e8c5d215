"""The subset of the ``astra`` namespace the reference touches, backed by the sm_100a engine.

``src/white_projection.py:8`` gets the name ``astra`` through ``from astra_proxy import *``, and
``teeth_recon/experiments.py:103-192`` carries a verbatim copy of astra_proxy that drives ASTRA's
object manager (``data2d.create/get/delete``, ``astra_dict``, ``algorithm.create/run/delete``).  Both
forms work against this shim.  Geometry dicts stay plain picklable dicts
(``teeth_recon/experiments.py:284-285``); the projector id is an int into a process-wide registry that
is never freed, like the reference's projector (SURVEY 8b "ownership").
"""
import itertools

import numpy as np

from .projector import Projector

_ids = itertools.count(1)
_projectors = {}
_data = {}
_algs = {}


def create_proj_geom(intype, *args):
    """astra.create_proj_geom('parallel', det_width, det_count, angles)  (src/white_projection.py:41-47)."""
    if intype != "parallel":
        raise NotImplementedError("only 'parallel' 2D geometries are on the hot path (got %r)" % (intype,))
    det_width, det_count, angles = args
    return {"type": "parallel", "DetectorWidth": det_width, "DetectorCount": det_count,
            "ProjectionAngles": np.asarray(angles)}


def create_vol_geom(*varargin):
    """astra.create_vol_geom(rows, cols) or create_vol_geom((rows, cols)) or create_vol_geom(n)
    (src/white_projection.py:60; teeth_recon/experiments.py:234,822)."""
    if len(varargin) == 1 and hasattr(varargin[0], "__len__"):
        varargin = tuple(varargin[0])
    if len(varargin) == 1:
        varargin = (varargin[0], varargin[0])
    if len(varargin) != 2:
        raise NotImplementedError("only 2D volume geometries are on the hot path")
    rows, cols = int(varargin[0]), int(varargin[1])
    return {"GridRowCount": rows, "GridColCount": cols,
            "option": {"WindowMinX": -cols / 2.0, "WindowMaxX": cols / 2.0,
                       "WindowMinY": -rows / 2.0, "WindowMaxY": rows / 2.0}}


def _geom_key(pg, vg):
    ang = np.asarray(pg["ProjectionAngles"], dtype=np.float64).astype(np.float32)
    return (int(vg["GridRowCount"]), int(vg["GridColCount"]), int(pg["DetectorCount"]),
            float(pg["DetectorWidth"]), ang.tobytes())


def create_projector(proj_type, proj_geom, vol_geom):
    """astra.create_projector('linear', pg, vg) -> id  (src/white_projection.py:61-63)."""
    if proj_type not in ("linear",):
        raise NotImplementedError("only the 'linear' (Joseph) projector is implemented (got %r)" % (proj_type,))
    pid = next(_ids)
    _projectors[pid] = Projector(vol_geom["GridRowCount"], vol_geom["GridColCount"], proj_geom["DetectorCount"],
                                 proj_geom["ProjectionAngles"], proj_geom["DetectorWidth"])
    _projectors[pid]._key = _geom_key(proj_geom, vol_geom)
    return pid


def projector_for(pg, vg, proj_id=None):
    """The engine object behind a projector id; checked against (pg, vg) like ASTRA checks shapes."""
    key = _geom_key(pg, vg)
    p = _projectors.get(proj_id)
    if p is not None and p._key == key:
        return p
    for q in _projectors.values():        # same geometry created elsewhere (e.g. unpickled dicts)
        if q._key == key:
            return q
    pid = create_projector("linear", pg, vg)
    return _projectors[pid]


# ---- object manager emulation (only what teeth_recon/experiments.py:103-192 / src/astra_proxy.py use) ----
class _Data2D(object):
    @staticmethod
    def create(datatype, geometry, data=None):
        if datatype not in ("-vol", "-sino"):
            raise ValueError("unknown data type %r" % (datatype,))
        if datatype == "-vol":
            shape = (geometry["GridRowCount"], geometry["GridColCount"])
        else:
            shape = (len(geometry["ProjectionAngles"]), geometry["DetectorCount"])
        if data is None:
            arr = np.zeros(shape, dtype=np.float32)
        else:
            arr = np.array(data, dtype=np.float32)          # ASTRA copies + casts to float32
            if arr.shape != tuple(shape):
                raise ValueError("The dimensions of the data do not match those specified in the geometry: "
                                 "%s != %s" % (arr.shape, tuple(shape)))
        i = next(_ids)
        _data[i] = {"type": datatype, "geom": geometry, "data": arr}
        return i

    @staticmethod
    def get(i):
        return _data[i]["data"].copy()

    @staticmethod
    def store(i, data):
        _data[i]["data"][...] = data

    @staticmethod
    def delete(ids):
        for i in (ids if hasattr(ids, "__len__") else [ids]):
            _data.pop(i, None)


class _Algorithm(object):
    @staticmethod
    def create(cfg):
        if cfg["type"] not in ("FP", "BP", "SIRT", "FBP"):
            raise NotImplementedError("algorithm %r is not on the hot path" % (cfg["type"],))
        i = next(_ids)
        _algs[i] = dict(cfg)
        return i

    @staticmethod
    def run(i, iterations=1):
        cfg = _algs[i]
        kind = cfg["type"]
        sino = _data[cfg["ProjectionDataId"]]
        vol = _data[cfg["VolumeDataId"] if kind == "FP" else cfg["ReconstructionDataId"]]
        proj = projector_for(sino["geom"], vol["geom"], cfg.get("ProjectorId"))
        if kind == "FP":
            sino["data"][...] = proj.fp(vol["data"]).cpu().numpy()
        elif kind == "BP":
            vol["data"][...] = proj.bp(sino["data"]).cpu().numpy()
        elif kind == "SIRT":
            # ASTRA continues from the current reconstruction; the reference always starts from zeros
            vol["data"][...] = proj.sirt(sino["data"], iterations).cpu().numpy()
        else:
            vol["data"][...] = proj.fbp(sino["data"]).cpu().numpy()

    @staticmethod
    def delete(ids):
        for i in (ids if hasattr(ids, "__len__") else [ids]):
            _algs.pop(i, None)


class _Projector(object):
    @staticmethod
    def delete(ids):
        for i in (ids if hasattr(ids, "__len__") else [ids]):
            _projectors.pop(i, None)


def astra_dict(intype):
    return {"type": intype}


data2d = _Data2D
algorithm = _Algorithm
projector = _Projector
