"""GPU: patch extraction of the input pipeline (data.py:119-142) on the hand-written group_knn, against the CPU
oracle (oracle/ref_net.group_knn, itself bit-identical to the unmodified reference)."""
import pytest
import torch

from oracle import ref_net
from tests.util import knn_gap_check

pytestmark = pytest.mark.gpu


def test_shape_to_patch_matches_reference_group_knn(pu3, cuda):
    g = torch.Generator().manual_seed(0)
    N, r, M, B = 5000, 4, 312, 6
    inp, lab = torch.rand(1, N, 3, generator=g), torch.rand(1, N * r, 3, generator=g)
    seeds = torch.randint(0, N, (B,), generator=g)
    ip, lp = pu3.patches.shape_to_patch(inp.to(cuda), lab.to(cuda), r, M, B, seed_idx=seeds)
    assert ip.shape == (B, M, 3) and lp.shape == (B, M * r, 3)
    rnd = inp[:, seeds, :]
    want_i, idx_i, _ = ref_net.group_knn(M, rnd, inp, unique=True, NCHW=False)
    want_l, idx_l, _ = ref_net.group_knn(M * r, rnd, lab, unique=True, NCHW=False)
    # the patches are exactly what group_knn returns for those seeds ...
    for got, pts, want_idx, k in ((ip, inp, idx_i, M), (lp, lab, idx_l, M * r)):
        knn_pts, idx, dist = pu3.operations.group_knn(k, rnd.to(cuda), pts.to(cuda), NCHW=False)
        assert torch.equal(got, knn_pts[0])
        assert torch.equal(got.cpu(), pts[0][idx[0].cpu()])                 # every returned point is a point of the cloud
        # ... and the neighbour sets agree with the oracle except at near-ties of the expanded-form distance
        knn_gap_check(rnd.transpose(1, 2), pts.transpose(1, 2), idx, dist, want_idx, k)
    assert torch.equal(ip[:, 0].cpu(), rnd[0])                             # rank 0 is the seed itself


def test_whole_shape_label_patch_size(pu3, cuda):
    """The largest live call of the loader: k = 312*16 = 4992 neighbours over an 80 000-point label shape."""
    g = torch.Generator().manual_seed(1)
    inp, lab = torch.rand(1, 5000, 3, generator=g).to(cuda), torch.rand(1, 80000, 3, generator=g).to(cuda)
    seeds = torch.tensor([0, 4999, 1234])
    ip, lp = pu3.patches.shape_to_patch(inp, lab, 16, 312, 3, seed_idx=seeds)
    assert lp.shape == (3, 4992, 3)
    rnd = inp[0, seeds.to(cuda)]
    d = torch.cdist(rnd.double().unsqueeze(0), lab.double())[0]            # (3, 80000)
    kth = d.topk(4992, dim=1, largest=False)[0][:, -1]
    got_d = (lp.double() - rnd.double().unsqueeze(1)).norm(dim=-1)
    assert bool((got_d.max(dim=1)[0] <= kth * (1 + 1e-5) + 1e-7).all())    # exactly the 4992 nearest (up to fp32 near-ties)
    assert bool((got_d[:, 1:] >= got_d[:, :-1] - 1e-6).all())              # ascending, like torch.topk(sorted=True)
