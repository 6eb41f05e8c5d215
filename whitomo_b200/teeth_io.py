"""On-disk formats either side of the hot path.

  read_data_from_hdf5   teeth_recon/teeth_io.py:47-54   h5 dataset 'sinogram'[angle, det, slice], 0.5 degree steps
  check_bounds          teeth_recon/teeth_recon.py:64-71  clip the line integrals to [minbound, maxbound] + masks
  normalize_data        teeth_recon/preproc_data.py:29-36  dark / flat-field correction and -log of raw counts
  slice_stack           the slice-major reader that feeds the C4 pipeline from such a cube
  savetxt / loadtxt     the `.nptxt` dumps of src/barrier_method_gogo.py:156-157 (np.savetxt of one element map)

The reference reads ONE slice per call (``h5f['sinogram'][:, :, slice_line]``: a gather with the slice index as the
fastest axis).  A stack job wants (S, A, D) with a slice contiguous, so ``slice_stack`` reads contiguous slice runs
``cube[:, :, s0:s1]`` -- whole rows of the file -- stages them in pinned host memory, uploads, and permutes on the
device; the permute is a plain strided copy in HBM (torch plumbing, no arithmetic).
"""
import numpy as np
import torch

_NATIVE = (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.int16), np.dtype(np.int32), np.dtype(np.uint8))
ANGLE_STEP_DEG = 0.5      # teeth_recon/teeth_io.py:52: "angular step 0.5 degree, measured from 0 to 200 degrees"


def angles_for(n_angles, step_deg=ANGLE_STEP_DEG):
    """projection angles in radians of a cube with ``n_angles`` rows (teeth_recon/teeth_io.py:53)"""
    return (np.arange(n_angles) * step_deg) * np.pi / 180.0


def open_cube(in_filename, dataset="sinogram"):
    """h5py dataset [angle, det, slice] (kept open by the returned file object: ``f, cube = open_cube(..)``)."""
    try:
        import h5py
    except ImportError as e:            # not in every image; the reader itself takes any array-like cube
        raise ImportError("reading %s needs h5py; pass any [angle, det, slice] array-like (e.g. a numpy memmap) "
                          "to slice_stack() instead" % in_filename) from e
    f = h5py.File(in_filename, "r")
    return f, f[dataset]


def read_data_from_hdf5(in_filename, slice_line):
    """{'data': (A, D) array of slice ``slice_line``, 'angles': radians} -- teeth_recon/teeth_io.py:47-54"""
    f, cube = open_cube(in_filename)
    try:
        data = np.asarray(cube[:, :, slice_line])
    finally:
        f.close()
    return {"data": data, "angles": angles_for(data.shape[0])}


def slice_stack(cube, slices, device, run=16, dtype=torch.float32):
    """(len(slices), A, D) ``dtype`` tensor on ``device`` from an [angle, det, slice] array-like ``cube``.

    ``slices`` is a range, slice or sequence of slice indices.  Consecutive indices are read as one run of at most
    ``run`` slices (``cube[:, :, s0:s1]``, contiguous rows of the file), copied through a pinned staging buffer and
    permuted to slice-major on the device."""
    A, D, S = cube.shape
    if isinstance(slices, slice):
        slices = range(*slices.indices(S))
    idx = [int(s) for s in slices]
    if any(s < 0 or s >= S for s in idx):
        raise IndexError("slice index out of range [0, %d)" % S)
    device = torch.device(device)
    out = torch.empty((len(idx), A, D), dtype=dtype, device=device)
    run = max(1, int(run))
    stage = None
    pos = 0
    while pos < len(idx):
        n = 1
        while pos + n < len(idx) and n < run and idx[pos + n] == idx[pos] + n:
            n += 1
        block = np.ascontiguousarray(cube[:, :, idx[pos]:idx[pos] + n])          # (A, D, n), one strided file read
        if block.dtype not in _NATIVE:
            block = block.astype(np.float32)                                     # e.g. uint16 detector counts
        tb = torch.from_numpy(block)
        if device.type == "cuda":
            if stage is None or stage.dtype != tb.dtype:
                stage = torch.empty((A, D, run), dtype=tb.dtype).pin_memory()
            stage[:, :, :n].copy_(tb)
            out[pos:pos + n].copy_(stage[:, :, :n].to(device, non_blocking=True).permute(2, 0, 1))
            torch.cuda.current_stream(device).synchronize()     # the staging buffer is reused by the next run
        else:
            out[pos:pos + n].copy_(tb.permute(2, 0, 1))
        pos += n
    return out


def check_bounds(r, maxbound, minbound):
    """clip ``r`` in place to [minbound, maxbound]; returns (r, above, below) -- teeth_recon/teeth_recon.py:64-71.
    Works on numpy arrays and on (device) tensors."""
    M = r > maxbound
    m = r < minbound
    r[M] = maxbound
    r[m] = minbound
    return r, M, m


def normalize_data(data, dark, empty, pix_size):
    """(line integrals, I0) from raw counts -- teeth_recon/preproc_data.py:29-36:  I0 = mean(empty - mean(dark));
    counts below the dark level are raised to it (in place, like the reference);
    r = (log(I0 + 1e-3) - log(data - mean(dark) + 1e-3)) / pix_size.  numpy arrays or (device) tensors."""
    if isinstance(data, torch.Tensor):
        avg_dark = torch.as_tensor(dark, dtype=torch.float64).mean().item()
        i0 = (torch.as_tensor(empty, dtype=torch.float64) - avg_dark).mean().item()
        data.clamp_(min=avg_dark)
        return (float(np.log(i0 + 1e-3)) - torch.log(data - avg_dark + 1e-3)) / pix_size, i0
    avg_dark = np.mean(dark)
    i0 = np.mean(empty - avg_dark)
    data[data < avg_dark] = avg_dark
    return (np.log(i0 + 1e-3) - np.log(data - avg_dark + 1e-3)) / pix_size, i0


def savetxt(path, x):
    """np.savetxt dump of one 2-D map (src/barrier_method_gogo.py:156-157); tensors are brought to the host first"""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    np.savetxt(path, np.asarray(x))


def loadtxt(path, device=None):
    """inverse of savetxt; float64 numpy array, or a float32 tensor on ``device``"""
    a = np.loadtxt(path)
    return a if device is None else torch.from_numpy(a.astype(np.float32)).to(device)
