"""Mirror pairing (calls with one or two images of >= 768^2 pixels run each ray together with its point reflection
through the V = 2 / V = 4 kernels) against the direct batched kernels, which tests/test_gpu_parity.py validates
against the oracle.  Every fused epilogue is covered: plain, SIRT, white-beam (K = 1, 2), MAR.  Needs a B200."""
import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
N, NA = 800, 96          # smallest square size class that takes the mirror path (>= 768^2), odd detector count


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from whitomo_b200.projector import Projector
    return Projector(N, N, int(np.sqrt(2) * N + 1), np.linspace(0, np.pi, NA), positions="exact")


def _imgs(n, seed):
    from whitomo_b200 import phantoms
    rng = np.random.default_rng(seed)
    base = phantoms.two_discs(N)
    return torch.from_numpy(np.stack([(base[i % 2] * rng.uniform(0.3, 1.0) + 0.02 * rng.random((N, N))).astype(np.float32)
                                      for i in range(n)])).cuda()


def test_plain_fp_bp(P):
    x = _imgs(3, 0)
    full = P.fp(x)
    for n in (1, 2):
        assert rel_l2(P.fp(x[:n])[0].cpu().numpy(), full[0].cpu().numpy()) < 1e-5
    y = torch.rand((3,) + P.sino_shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    fullb = P.bp(y)
    for n in (1, 2):
        assert rel_l2(P.bp(y[:n], scale=-2.0)[n - 1].cpu().numpy(), -2.0 * fullb[n - 1].cpu().numpy()) < 1e-5
    # odd row count: the middle row is its own mirror row
    from whitomo_b200.projector import Projector
    Q = Projector(801, 799, 1133, np.linspace(0, np.pi, 16), positions="exact")
    yq = torch.rand((3,) + Q.sino_shape, device="cuda")
    xq = torch.rand((3, 801, 799), device="cuda")
    assert rel_l2(Q.bp(yq[:1])[0].cpu().numpy(), Q.bp(yq)[0].cpu().numpy()) < 1e-5
    assert rel_l2(Q.fp(xq[:1])[0].cpu().numpy(), Q.fp(xq)[0].cpu().numpy()) < 1e-5


def test_sirt(P):
    p = P.fp(_imgs(3, 2))
    ref = P.sirt(p, 6)
    assert rel_l2(P.sirt(p[0], 6).cpu().numpy(), ref[0].cpu().numpy()) < 1e-4
    assert rel_l2(P.sirt(p[:2], 6)[1].cpu().numpy(), ref[1].cpu().numpy()) < 1e-4


@pytest.mark.parametrize("K", [1, 2])
def test_white_beam_epilogue(P, K):
    from whitomo_b200.projector import Spectrum
    from whitomo_b200 import phantoms
    grid, source, ea, pix = phantoms.synth_spectrum(64, N)
    sp = Spectrum(source / np.sum(source * grid[:, 1]), grid[:, 1], ea[:K], pix)
    S = 3 if K == 1 else 2
    conc = (_imgs(S * K, 3) * 0.3).reshape(S, K, N, N)
    meas = torch.rand((S,) + P.sino_shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    inten, qmu, loss = P.white_fp(sp, conc, meas=meas)                 # S*K >= 3 images: direct kernels
    for s_take in ((1,) if K == 2 else (1, 2)):                         # 1 or 2 images: mirror pairing
        i2, q2, l2 = P.white_fp(sp, conc[:s_take], meas=meas[:s_take])
        assert rel_l2(i2.cpu().numpy(), inten[:s_take].cpu().numpy()) < 1e-5
        assert rel_l2(q2.cpu().numpy(), qmu[:s_take].cpu().numpy()) < 1e-4
        assert torch.allclose(l2, loss[:s_take], rtol=1e-5)


def test_mar_epilogue(P):
    x = _imgs(3, 5)
    gen = torch.Generator(device="cuda").manual_seed(6)
    b = torch.rand((3,) + P.sino_shape, device="cuda", generator=gen) * 50
    mask = (torch.rand((3,) + P.sino_shape, device="cuda", generator=gen) < 0.1).to(torch.uint8)
    inv = torch.tensor([1e-4, 2e-4, 3e-4], device="cuda")
    u, pr, st = P.mar_fp(x, b, mask, inv, 0.2, -10.0, want_proj=True)
    u1, pr1, st1 = P.mar_fp(x[:1], b[:1], mask[:1], inv[:1], 0.2, -10.0, want_proj=True)
    assert rel_l2(pr1.cpu().numpy(), pr[:1].cpu().numpy()) < 1e-5
    assert rel_l2(u1.cpu().numpy(), u[:1].cpu().numpy()) < 1e-4
    fin = torch.isfinite(st[0])
    assert torch.equal(torch.isfinite(st1[0]), fin)
    assert torch.allclose(st1[0][fin], st[0][fin], rtol=1e-4)
