"""On-disk formats either side of the hot path: the [angle, det, slice] cube reader, .nptxt dumps and
the real-data geometry of teeth_recon/teeth_recon.py:73-138 (volume side = detector count, 0.5 degree steps over
200 degrees) through the engine against the oracle."""
import os

import numpy as np
import pytest

import oracle
from conftest import rel_l2

torch = pytest.importorskip("torch")
from whitomo_b200 import teeth_io


def _cube(tmp_path, A=9, D=6, S=40, dtype=np.uint16):
    cube = np.random.default_rng(3).integers(0, 60000, (A, D, S)).astype(dtype)
    f = os.path.join(str(tmp_path), "cube.dat")
    cube.tofile(f)
    return cube, np.memmap(f, dtype=dtype, mode="r", shape=(A, D, S))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_slice_stack_matches_per_slice_reads(tmp_path, dtype):
    cube, mm = _cube(tmp_path, dtype=dtype)
    want = [0, 1, 2, 3, 9, 10, 30, 5]
    out = teeth_io.slice_stack(mm, want, "cpu", run=3)
    # the reference's access pattern: h5f['sinogram'][:, :, slice_line]  (teeth_recon/teeth_io.py:49)
    ref = np.stack([cube[:, :, s] for s in want]).astype(np.float32)
    assert out.dtype == torch.float32 and np.array_equal(out.numpy(), ref)
    assert teeth_io.slice_stack(mm, slice(4, None, 2), "cpu").shape == (18, 9, 6)
    assert teeth_io.slice_stack(mm, [], "cpu").shape == (0, 9, 6)
    with pytest.raises(IndexError):
        teeth_io.slice_stack(mm, [40], "cpu")


def test_angles_bounds_and_nptxt(tmp_path):
    a = teeth_io.angles_for(401)
    assert a[0] == 0.0 and np.isclose(a[-1], 200.0 * np.pi / 180.0)
    r = np.array([[-1.0, 0.5], [3.5, 2.9]])
    rr, M, m = teeth_io.check_bounds(r, 2.9, 0)
    assert rr is r and np.array_equal(r, [[0.0, 0.5], [2.9, 2.9]])
    assert np.array_equal(M, [[False, False], [True, False]]) and np.array_equal(m, [[True, False], [False, False]])
    t = torch.tensor([[-1.0, 0.5], [3.5, 2.9]])
    tt, _, _ = teeth_io.check_bounds(t, 2.9, 0)
    assert np.array_equal(tt.numpy(), r.astype(np.float32))
    x = np.random.default_rng(0).random((5, 7))
    p = os.path.join(str(tmp_path), "x0_0001.nptxt")
    teeth_io.savetxt(p, torch.from_numpy(x))
    assert np.array_equal(teeth_io.loadtxt(p), np.loadtxt(p)) and np.allclose(teeth_io.loadtxt(p), x, rtol=0, atol=1e-15)


def test_normalize_data_numpy_and_tensor_agree():
    rng = np.random.default_rng(2)
    dark = rng.uniform(90, 110, (3, 8, 8))
    empty = rng.uniform(5000, 5200, (3, 8, 8))
    data = rng.uniform(50, 5000, (5, 8, 8))
    avg_dark = dark.mean()
    i0_ref = (empty - avg_dark).mean()
    ref = (np.log(i0_ref + 1e-3) - np.log(np.maximum(data, avg_dark) - avg_dark + 1e-3)) / 2.0   # preproc_data.py:29-36
    r_np, i0 = teeth_io.normalize_data(data.copy(), dark, empty, 2.0)
    assert np.allclose(r_np, ref, rtol=0, atol=1e-12) and np.isclose(i0, i0_ref)
    r_t, i0_t = teeth_io.normalize_data(torch.from_numpy(data.copy()), dark, empty, 2.0)
    assert np.allclose(r_t.numpy(), ref, rtol=0, atol=1e-9) and np.isclose(i0_t, i0_ref)
    assert r_np.min() >= -1e-12 + np.log(i0_ref + 1e-3) / 2.0 - np.log(5000 - avg_dark + 1e-3) / 2.0


def test_hdf5_reader_fails_loudly_without_h5py(tmp_path):
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            teeth_io.read_data_from_hdf5(os.path.join(str(tmp_path), "zub_clear.h5"), 820)
    else:
        import h5py
        cube = np.random.default_rng(1).random((11, 8, 5)).astype(np.float32)
        f = os.path.join(str(tmp_path), "zub.h5")
        with h5py.File(f, "w") as h:
            h["sinogram"] = cube
        d = teeth_io.read_data_from_hdf5(f, 3)
        assert np.array_equal(d["data"], cube[:, :, 3]) and d["angles"].shape == (11,)


@pytest.mark.gpu
def test_real_data_geometry_through_the_engine(tmp_path):
    """cube -> slice-major device stack -> check_bounds -> FBP / FP / BP with rec_pix = detector count and 200 degrees
    of 0.5 degree steps (teeth_recon/teeth_recon.py:114-138), against the oracle on the same slices."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from whitomo_b200.projector import Projector
    D, A, S = 96, 401, 6
    ang = teeth_io.angles_for(A)
    g = oracle.Geometry(D, D, D, ang)
    yy, xx = np.mgrid[0:D, 0:D]
    cube = np.empty((A, D, S), np.float32)
    for s in range(S):
        ph = (((xx - 40 - 2 * s) ** 2 + (yy - 50) ** 2) <= (12 + s) ** 2).astype(np.float32) * 0.12
        cube[:, :, s] = oracle.fp(g, ph)
    f = os.path.join(str(tmp_path), "cube.dat")
    cube.tofile(f)
    mm = np.memmap(f, dtype=np.float32, mode="r", shape=(A, D, S))
    stack = teeth_io.slice_stack(mm, range(1, 5), "cuda", run=3)
    assert stack.is_cuda and np.array_equal(stack.cpu().numpy(), np.stack([cube[:, :, s] for s in range(1, 5)]))
    bound = 2.9
    stack, M, m = teeth_io.check_bounds(stack, bound, 0)
    P = Projector(D, D, D, ang)
    rec = P.fbp(stack).cpu().numpy()
    back = P.bp(stack).cpu().numpy()
    for i, s in enumerate(range(1, 5)):
        sino = np.clip(cube[:, :, s], 0, bound)
        assert rel_l2(rec[i], oracle.fbp(g, sino)) < 1e-4
        assert rel_l2(back[i], oracle.bp(g, sino)) < 1e-4
    x = torch.from_numpy(rec[:2].copy()).cuda()
    assert rel_l2(P.fp(x).cpu().numpy()[1], oracle.fp(g, rec[1])) < 1e-4
