"""The cluster-split paths of K1 / K2 (small geometries: lines / angle blocks of a unit of work spread over the CTAs of a
thread-block cluster, partial sums reduced through distributed shared memory) against the same kernels run unsplit.
The choice is read from the environment once per process, so the two runs are separate processes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from whitomo_b200.projector import Projector
for n, na, imgs in ((256, 180, 1), (200, 97, 3), (384, 270, 2), (512, 360, 1)):
    d = int(np.sqrt(2) * n + 1)
    P = Projector(n, n - 8, d, np.linspace(0, np.pi, na, endpoint=False))
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.rand((imgs, n, n - 8), device="cuda", generator=g)
    y = torch.rand((imgs, na, d), device="cuda", generator=g)
    s = P.sirt(P.fp(x), 5)
    torch.save({"fp": P.fp(x).cpu(), "bp": P.bp(y).cpu(), "bp_half": P.bp(y, angle_range=(3, na // 2)).cpu(), "sirt": s.cpu()}, sys.argv[1] + "_%%d.pt" %% n)
print("done")
"""


def _run(tmp, tag, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    prefix = os.path.join(tmp, tag)
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT, prefix], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return prefix


@pytest.mark.gpu
def test_cluster_split_matches_unsplit_kernels(tmp_path):
    import torch
    a = _run(str(tmp_path), "split", {})
    b = _run(str(tmp_path), "plain", {"WT_NO_SPLIT": "1"})
    for n in (256, 200, 384, 512):
        sa, sb = torch.load("%s_%d.pt" % (a, n)), torch.load("%s_%d.pt" % (b, n))
        for k in ("fp", "bp", "bp_half", "sirt"):
            num = float(torch.linalg.vector_norm((sa[k] - sb[k]).double()))
            den = float(torch.linalg.vector_norm(sb[k].double()))
            # same taps, same weights; only the order of the partial sums differs
            assert num <= 2e-6 * den, (n, k, num / den)
            assert torch.isfinite(sa[k]).all()


@pytest.mark.gpu
def test_cluster_split_is_run_to_run_identical(tmp_path):
    import torch
    a = _run(str(tmp_path), "r1", {})
    b = _run(str(tmp_path), "r2", {})
    for n in (256, 200, 384, 512):
        sa, sb = torch.load("%s_%d.pt" % (a, n)), torch.load("%s_%d.pt" % (b, n))
        for k in ("fp", "bp", "bp_half", "sirt"):
            assert torch.equal(sa[k], sb[k]), (n, k)
