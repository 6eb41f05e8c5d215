"""Device-tensor interface to the sm_100a kernels: one ``Projector`` per (volume geometry, projection
geometry) pair -- the object ``astra.create_projector('linear', pg, vg)`` stands for in the reference
(src/white_projection.py:61-63).  PyTorch is used for device memory and streams only; all arithmetic
happens in libwhitomo_b200.so, reached through the C ABI in include/whitomo_b200.h.
"""
import ctypes

import numpy as np
import torch

from . import _capi


def _require_cuda():
    if not torch.cuda.is_available():
        raise _capi.EngineError("whitomo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_own_device(method):
    """run a Projector method with the projector's device current (streams and launches are per device)"""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        if torch.cuda.current_device() == self.device.index:
            return method(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapper


class Spectrum(object):
    """Device tables of the polychromatic model (constructor state of WhiteProjection,
    src/white_projection.py:23-73).  ``source`` must already be normalised by its integral."""

    def __init__(self, source, widths, absorptions, pixel_size, device=None):
        """``device``: the GPU the tables live on (default: the current one); use the Projector's device."""
        _require_cuda()
        self.lib = _capi.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        source = np.ascontiguousarray(source, dtype=np.float64)
        widths = np.ascontiguousarray(widths, dtype=np.float64)
        absorptions = np.ascontiguousarray(absorptions, dtype=np.float64)
        self.K, self.E = absorptions.shape
        assert source.shape == (self.E,) and widths.shape == (self.E,)
        # host copies: what a checkpoint of a reconstruction depends on (solvers._checkpoint_key)
        self.source, self.widths, self.absorptions, self.pixel_size = source, widths, absorptions, float(pixel_size)
        h = ctypes.c_void_p()
        dp = ctypes.POINTER(ctypes.c_double)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.wt_spectrum_create(self.K, self.E, source.ctypes.data_as(dp), widths.ctypes.data_as(dp),
                                                    absorptions.ctypes.data_as(dp), float(pixel_size), ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.wt_spectrum_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Projector(object):
    """``positions``: "astra" (default; the reference's float32 running-sum ray positions, include/whitomo_b200.h
    WT_POS_ASTRA) or "exact" (closed-form positions, mirror-paired single-image kernels); None reads the environment
    variable WT_POSITIONS."""

    def __init__(self, rows, cols, ndet, angles, det_width=1.0, device=None, positions=None):
        _require_cuda()
        self.lib = _capi.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.rows, self.cols, self.ndet = int(rows), int(cols), int(ndet)
        self.det_width = float(det_width)
        self.angles = np.ascontiguousarray(np.asarray(angles, dtype=np.float64).astype(np.float32))
        self.nangles = int(self.angles.shape[0])
        if positions is None:
            import os
            positions = os.environ.get("WT_POSITIONS", "astra") or "astra"
        if positions not in ("astra", "exact"):
            raise ValueError("positions must be 'astra' or 'exact'")
        self.positions = positions
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.wt_geom_create_ex(self.rows, self.cols, self.ndet, self.det_width,
                                                   self.angles.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                                   self.nangles, _capi.POS_EXACT if positions == "exact" else _capi.POS_ASTRA,
                                                   ctypes.byref(h)))
        self.handle = h
        self._ws = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.wt_geom_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- shapes -------------------------------------------------------------------------------
    @property
    def vol_shape(self):
        return (self.rows, self.cols)

    @property
    def sino_shape(self):
        return (self.nangles, self.ndet)

    def _workspace(self, nbytes):
        """scratch for the packed copies and partial sums, one buffer per CUDA stream: calls issued on different
        streams may overlap on the device, calls on one stream are ordered (so is the allocator's reuse on growth)"""
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            self._ws.pop(key, None)
            ws = self._ws[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return ws

    def _as_images(self, t, shape):
        """accept (..., *shape) float32 CUDA tensors; return (contiguous (n, *shape) view, leading dims)"""
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t, dtype=np.float32))
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        if t.dtype != torch.float32:
            t = t.float()
        if tuple(t.shape[-2:]) != tuple(shape):
            raise ValueError("expected trailing shape %s, got %s" % (shape, tuple(t.shape)))
        lead = tuple(t.shape[:-2])
        return t.reshape((-1,) + tuple(shape)).contiguous(), lead

    # ---- operators ------------------------------------------------------------------------------
    @_on_own_device
    def fp(self, vol, out=None, angle_range=None):
        """p = W x.  vol (..., rows, cols) -> (..., nangles, ndet) float32 on the device.  With angle_range only
        those sinogram rows are computed (the others are zero in a fresh output)."""
        v, lead = self._as_images(vol, self.vol_shape)
        n = v.shape[0]
        a0, a1 = (0, self.nangles) if angle_range is None else angle_range
        if out is None:
            alloc = torch.empty if angle_range is None else torch.zeros
            out = alloc((n,) + self.sino_shape, dtype=torch.float32, device=self.device)
        if n == 0:                                   # empty batch: nothing to launch
            return out.reshape(lead + self.sino_shape)
        nbytes = self.lib.wt_workspace_bytes(self.handle, n)
        ws = self._workspace(nbytes)
        _capi.check(self.lib.wt_fp(self.handle, _ptr(v), _ptr(out), n, int(a0), int(a1), _ptr(ws), nbytes, _stream()))
        return out.reshape(lead + self.sino_shape)

    @_on_own_device
    def bp(self, sino, out=None, angle_range=None, scale=1.0):
        """g = scale * W^T q (exact transpose).  sino (..., nangles, ndet) -> (..., rows, cols)."""
        s, lead = self._as_images(sino, self.sino_shape)
        n = s.shape[0]
        if out is None:
            out = torch.empty((n,) + self.vol_shape, dtype=torch.float32, device=self.device)
        a0, a1 = (0, self.nangles) if angle_range is None else angle_range
        if n == 0:
            return out.reshape(lead + self.vol_shape)
        nbytes = self.lib.wt_workspace_bytes(self.handle, n)
        ws = self._workspace(nbytes)
        _capi.check(self.lib.wt_bp(self.handle, _ptr(s), _ptr(out), n, int(a0), int(a1), float(scale), _ptr(ws), nbytes,
                                   _stream()))
        return out.reshape(lead + self.vol_shape)

    def bp_row_blocks(self, nimages, nblocks):
        """Row blocks for `bp_rows`: list of (row_begin, row_end, [written row ranges]).  Blocks are aligned to the
        backprojector's tile height; with mirror-paired single images (positions="exact") a block also writes its
        point reflection, so it lists two row ranges."""
        rp, al, mir = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _capi.check(self.lib.wt_bp_row_info(self.handle, int(nimages), ctypes.byref(rp), ctypes.byref(al), ctypes.byref(mir)))
        rp, al, mir = rp.value, al.value, mir.value
        per = -(-rp // max(1, nblocks))
        per = -(-per // al) * al
        out = []
        for b0 in range(0, rp, per):
            b1 = min(rp, b0 + per)
            written = [(b0, b1)]
            if mir:
                lo, hi = self.rows - b1, self.rows - b0
                lo = max(lo, b1) if lo < b1 else lo       # the middle row of an odd image is written once
                if hi > lo:
                    written.append((lo, hi))
            out.append((b0, b1, written))
        return out

    @_on_own_device
    def bp_rows(self, sino, out, row_begin, row_end, angle_range=None, scale=1.0, reuse_packed=False):
        """rows [row_begin, row_end) of scale * W^T q written into `out` (n, rows, cols) (wt_bp_rows); the other rows
        of `out` are left untouched.  Blocks come from `bp_row_blocks`."""
        s, lead = self._as_images(sino, self.sino_shape)
        n = s.shape[0]
        a0, a1 = (0, self.nangles) if angle_range is None else angle_range
        nbytes = self.lib.wt_workspace_bytes(self.handle, n)
        ws = self._workspace(nbytes)
        _capi.check(self.lib.wt_bp_rows(self.handle, _ptr(s), _ptr(out), n, int(a0), int(a1), int(row_begin), int(row_end),
                                        int(bool(reuse_packed)), float(scale), _ptr(ws), nbytes, _stream()))
        return out

    @_on_own_device
    def sirt(self, sino, n_iters=100):
        """ASTRA-style SIRT from x0 = 0; sino (..., nangles, ndet) -> (..., rows, cols)."""
        s, lead = self._as_images(sino, self.sino_shape)
        n = s.shape[0]
        out = torch.empty((n,) + self.vol_shape, dtype=torch.float32, device=self.device)
        nbytes = self.lib.wt_sirt_workspace_bytes(self.handle, n)
        ws = self._workspace(nbytes)
        _capi.check(self.lib.wt_sirt(self.handle, _ptr(s), _ptr(out), n, int(n_iters), _ptr(ws), nbytes, _stream()))
        return out.reshape(lead + self.vol_shape)

    @_on_own_device
    def ramp_filter(self, sino):
        """Fourier ramp of ASTRA's CPU FBP: zero-pad to pow2 >= 2D, multiply bin i by 2 i / zp."""
        s, lead = self._as_images(sino, self.sino_shape)
        zp = 1
        while zp < 2 * self.ndet:
            zp <<= 1
        f = torch.fft.rfft(s, n=zp, dim=-1)
        ramp = (2.0 * torch.arange(zp // 2 + 1, device=self.device, dtype=torch.float32)) / zp
        out = torch.fft.irfft(f * ramp, n=zp, dim=-1)[..., :self.ndet].contiguous()
        return out.reshape(lead + self.sino_shape)

    def fbp(self, sino):
        """x = pi/(2A) * W^T ramp(p)  (cpu_fbp, src/astra_proxy.py:87-102)."""
        return self.bp(self.ramp_filter(sino), scale=np.pi / (2.0 * self.nangles))

    @_on_own_device
    def white_fp(self, spectrum, conc, meas=None, want_intensity=True, want_qmu=True, want_loss=True,
                 angle_range=None):
        """Fused white-beam forward model + residual (wt_white_fp).  conc (S, K, rows, cols) or (K, rows, cols).
        Returns (intensity (S,A,D) | None, qmu (S,K,A,D) | None, loss (S,) float64 | None)."""
        if spectrum.device != self.device:
            raise ValueError("the Spectrum lives on %s, the Projector on %s" % (spectrum.device, self.device))
        c, lead = self._as_images(conc, self.vol_shape)
        K = spectrum.K
        if c.shape[0] % K:
            raise ValueError("conc must hold a multiple of K=%d images" % K)
        S = c.shape[0] // K
        m = None
        if meas is not None:
            m, _ = self._as_images(meas, self.sino_shape)
            if m.shape[0] != S:
                raise ValueError("meas must hold one sinogram per slice")
        a0, a1 = (0, self.nangles) if angle_range is None else angle_range
        alloc = torch.empty if angle_range is None else torch.zeros
        inten = alloc((S,) + self.sino_shape, dtype=torch.float32, device=self.device) if want_intensity else None
        qmu = alloc((S, K) + self.sino_shape, dtype=torch.float32, device=self.device) if want_qmu else None
        loss = torch.empty((S,), dtype=torch.float64, device=self.device) if want_loss else None
        nbytes = self.lib.wt_workspace_bytes(self.handle, S * K)
        ws = self._workspace(nbytes)
        _capi.check(self.lib.wt_white_fp(self.handle, spectrum.handle, _ptr(c), _ptr(m), _ptr(inten), _ptr(qmu),
                                         _ptr(loss), S, int(a0), int(a1), _ptr(ws), nbytes, _stream()))
        return inten, qmu, loss

    def step_supported(self, spectrum, nslices):
        """whether `white_barrier_step` (the three-launch fused iteration) is available for this problem"""
        mir = ctypes.c_int()
        _capi.check(self.lib.wt_bp_row_info(self.handle, int(nslices * spectrum.K), None, None, ctypes.byref(mir)))
        return spectrum.K <= 2 and not mir.value

    def new_step_state(self, spectrum, nslices):
        """device scratch that carries the packed state of `white_barrier_step` between calls (one per working set)"""
        nbytes = int(self.lib.wt_step_workspace_bytes(self.handle, int(nslices), int(spectrum.K)))
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    @_on_own_device
    def white_barrier_step(self, spectrum, x, meas, state, init=False, alpha=0.0, beta_reg=1.0, t=0.0, reg_w=0.0, triv_w=0.0,
                           lower_only=False, mode=_capi.UPD_STEP_CLIP, elem_mask=0xFFFFFFFF, active=None,
                           want_loss=True, want_penalty=False, timed=False):
        """One fused inner iteration (wt_white_barrier_step): x (S, K, rows, cols) is stepped in place along
        -W^T(q mu) with the wt_barrier_update rule, then the white-beam forward model runs at the new x.  With
        ``init`` only the forward model runs (start of a run, or x was written from outside).  ``state`` comes from
        `new_step_state` and must be passed unchanged from an ``init`` call on.
        Returns (loss (S,) float64 = f at the returned x | None, penalty (S,) float64 | None)."""
        if x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device or x.dim() != 4:
            raise ValueError("x must be a contiguous float32 (S, K, rows, cols) tensor on %s" % self.device)
        if spectrum.device != self.device:
            raise ValueError("the Spectrum lives on %s, the Projector on %s" % (spectrum.device, self.device))
        S = x.shape[0]
        m = None
        if meas is not None:
            m, _ = self._as_images(meas, self.sino_shape)
            if m.shape[0] != S:
                raise ValueError("meas must hold one sinogram per slice")
        prm = _capi.UpdateParams(float(alpha), float(beta_reg), float(t), float(reg_w), float(triv_w),
                                 int(bool(lower_only)), int(mode), int(elem_mask) & 0xFFFFFFFF)
        act = None
        if active is not None:
            act = active.to(torch.uint8).contiguous()
            assert act.numel() == S
        loss = torch.empty((S,), dtype=torch.float64, device=self.device) if want_loss else None
        pen = torch.empty((S,), dtype=torch.float64, device=self.device) if (want_penalty and not init) else None
        _capi.check(self.lib.wt_white_barrier_step(self.handle, spectrum.handle, _ptr(x), _ptr(m), ctypes.byref(prm), _ptr(act),
                                                   _ptr(loss), _ptr(pen), S,
                                                   (_capi.STEP_INIT if init else 0) | (_capi.STEP_TIMED if timed else 0),
                                                   _ptr(state), state.numel(), _stream()))
        return loss, pen

    def step_times(self, max_steps=_capi.STEP_RING):
        """(bp_ms, fp_ms) lists of the most recent ``timed`` fused steps (device events recorded inside the call)"""
        n = min(int(max_steps), _capi.STEP_RING)
        b, f = (ctypes.c_float * n)(), (ctypes.c_float * n)()
        got = self.lib.wt_step_times(self.handle, n, b, f)
        if got < 0:
            _capi.check(got)
        return list(b[:got]), list(f[:got])

    @_on_own_device
    def mar_fp(self, vol, b, mask, inv_ngood, bt, b_hb, want_u=True, want_proj=False, want_stats=True):
        """Fused mono-MAR residual (wt_mar_fp).  vol (n, R, C); b (n, A, D); mask (n, A, D) uint8; inv_ngood (n,).
        Returns (u | None, proj | None, stats (n, 3) float64 | None): cost, hough phi sum, #infeasible metal rays."""
        v, _ = self._as_images(vol, self.vol_shape)
        n = v.shape[0]
        bb, _ = self._as_images(b, self.sino_shape)
        mm = mask.to(self.device, dtype=torch.uint8).reshape((n,) + self.sino_shape).contiguous()
        ig = inv_ngood.to(self.device, dtype=torch.float32).reshape(n).contiguous()
        u = torch.empty((n,) + self.sino_shape, dtype=torch.float32, device=self.device) if want_u else None
        pr = torch.empty((n,) + self.sino_shape, dtype=torch.float32, device=self.device) if want_proj else None
        st = torch.empty((n, 3), dtype=torch.float64, device=self.device) if want_stats else None
        nbytes = self.lib.wt_workspace_bytes(self.handle, n)
        ws = self._workspace(nbytes)
        _capi.check(self.lib.wt_mar_fp(self.handle, _ptr(v), _ptr(bb), _ptr(mm), _ptr(ig), float(bt), float(b_hb),
                                       _ptr(u), _ptr(pr), _ptr(st), n, _ptr(ws), nbytes, _stream()))
        return u, pr, st

    @_on_own_device
    def pair_norm2(self, conc, out=None):
        """Partial sums of sum_i (c0_i c1_i)^2 per slice (wt_pair_norm2): conc (S, 2, rows, cols) float32 ->
        (S, PAIR_NORM_BLOCKS) float64; ``.sum(-1).sqrt()`` is the ||c0 c1|| of src/white_rec.py:236-237.  ``out``: where
        to write them (a solver loop passes a row of its history buffer: one launch per iteration, nothing else)."""
        if conc.dtype != torch.float32 or not conc.is_contiguous() or conc.device != self.device or conc.dim() != 4 or conc.shape[1] != 2:
            raise ValueError("conc must be a contiguous float32 (S, 2, rows, cols) tensor on %s" % self.device)
        S = conc.shape[0]
        if out is None:
            out = torch.empty((S, _capi.PAIR_NORM_BLOCKS), dtype=torch.float64, device=self.device)
        elif out.dtype != torch.float64 or not out.is_contiguous() or out.numel() != S * _capi.PAIR_NORM_BLOCKS or out.device != self.device:
            raise ValueError("out must be a contiguous float64 tensor of S x %d values" % _capi.PAIR_NORM_BLOCKS)
        _capi.check(self.lib.wt_pair_norm2(_ptr(conc), S, conc.shape[-1] * conc.shape[-2], _ptr(out), _stream()))
        return out

    def barrier_update(self, x, grad, alpha, beta_reg, t, reg_w=0.0, triv_w=0.0, lower_only=False,
                       mode=_capi.UPD_STEP_CLIP, elem_mask=0xFFFFFFFF, active=None, want_penalty=False):
        """In-place fused K3 step (wt_barrier_update).  x, grad: (S, K, rows, cols) / (K, rows, cols) / (rows, cols)."""
        if x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device:
            raise ValueError("x must be a contiguous float32 tensor on %s" % self.device)
        g = grad
        if g is not None and not (g.dtype == torch.float32 and g.is_contiguous()):
            g = g.float().contiguous()
        if x.dim() == 2:
            S, K = 1, 1
        elif x.dim() == 3:
            S, K = 1, x.shape[0]
        else:
            S, K = x.shape[0], x.shape[1]
        npix = x.shape[-1] * x.shape[-2]
        prm = _capi.UpdateParams(float(alpha), float(beta_reg), float(t), float(reg_w), float(triv_w),
                                 int(bool(lower_only)), int(mode), int(elem_mask) & 0xFFFFFFFF)
        act = None
        if active is not None:
            act = active.to(torch.uint8).contiguous()
            assert act.numel() == S
        pen = torch.empty((S,), dtype=torch.float64, device=self.device) if want_penalty else None
        nbytes = max(int(self.lib.wt_workspace_bytes(self.handle, S * K)), S * 296 * 8)
        ws = self._workspace(nbytes)
        _capi.check(self.lib.wt_barrier_update(_ptr(x), _ptr(g), S, K, npix, ctypes.byref(prm), _ptr(act), _ptr(pen),
                                               _ptr(ws), nbytes, _stream()))
        return pen
