"""Behaviour ON the faces of the box (x == 0 or x == 1 exactly), where the reference's arithmetic produces signed
infinities: grad phi = -g/f (src/barrier_method.py:50) is -inf on both faces (1/(-0.0) for f = -x, -1/(+0.0) for
f = x - 1), so a pixel clipped onto a face is thrown to +inf by the step and clipped to 1 (box) -- the fused K3
kernel must reproduce exactly that, because whole trajectories depend on it.  Needs a B200: -m gpu."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _reference_step(x, g, alpha, beta, t, lower_only):
    """the reference's numpy statements for the box / morezero constraints, float64"""
    x = x.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        grad = g.astype(np.float64).copy()
        f, gg = -x, -np.ones_like(x)                       # conc_non_negative (src/barrier_method_gogo.py:87-89)
        grad += beta * (-gg / f) * t
        if not lower_only:
            f, gg = x - 1, np.ones_like(x)                 # conc_less_than_one (:92-94)
            grad += beta * (-gg / f) * t
        xn = x + (-alpha * grad)
        xn = np.clip(xn, 0, xn.max())
        if not lower_only:
            xn = np.clip(xn, xn.min(), 1)
    return xn


@pytest.mark.parametrize("lower_only", [False, True])
def test_faces_of_the_box(lower_only):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from whitomo_b200.projector import Projector
    P = Projector(8, 8, 12, np.linspace(0, np.pi, 4))
    rng = np.random.default_rng(0)
    K = 1 if lower_only else 2
    x = rng.uniform(0.1, 0.9, (1, K, 8, 8)).astype(np.float32)
    x[0, 0, 0, :4] = 0.0                                    # on the lower face
    if not lower_only:
        x[0, 1, 1, :4] = 1.0                                # on the upper face
    g = rng.normal(0, 0.3, x.shape).astype(np.float32)
    ref = _reference_step(x[0], g[0], 0.3, 1.0, 0.05, lower_only)
    xt = torch.from_numpy(x.copy()).cuda()
    pen = P.barrier_update(xt, torch.from_numpy(g).cuda(), 0.3, 1.0, 0.05, lower_only=lower_only, want_penalty=True)
    out = xt.cpu().numpy()[0].astype(np.float64)
    assert np.array_equal(np.isinf(out), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.allclose(out[fin], ref[fin], rtol=1e-5, atol=1e-7)
    if not lower_only:
        assert (out[0, 0, :4] == 1.0).all() and (out[1, 1, :4] == 1.0).all()      # thrown to +inf, clipped to 1
        assert np.isinf(pen.item())                                                # phi = inf on the face
    else:
        assert np.isinf(out[0, 0, :4]).all()                                       # morezero has no upper clip
