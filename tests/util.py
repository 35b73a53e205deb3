"""Shared checkers for the parity tests."""
import numpy as np
import torch


def bits_equal(a, b):
    """Bit-for-bit equality of two float arrays (distinguishes -0/+0, equal NaN payloads pass)."""
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def knn_gap_check(query, points, idx, dist, ref_idx, k, penalty=None, tol=4e-6):
    """Gap-aware comparison of kNN results.

    query (B,C,M), points (B,C,N) float32 CPU tensors; idx/dist the result under test, ref_idx the oracle's.
    Distances of the expanded form carry an absolute rounding error of about tol * (|q|^2 + |p|^2), and
    torch.topk leaves the order of ties unspecified, so indices must agree exactly only where the float64
    distance of the two candidates differs by more than that.  penalty: optional (B,1,N) float64 tensor added
    to the float64 distances (the duplicate penalty).
    Returns the number of positions where the indices differ (all of them justified by a near-tie)."""
    q = query.double().transpose(1, 2)  # B,M,C
    p = points.double().transpose(1, 2)  # B,N,C
    if p.size(0) != q.size(0):
        p = p.repeat_interleave(q.size(0) // p.size(0), dim=0)
    D = (q * q).sum(-1, keepdim=True) - 2 * q @ p.transpose(1, 2) + (p * p).sum(-1).unsqueeze(1)
    if penalty is not None:
        D = D + penalty
    scale = (q * q).sum(-1, keepdim=True) + (p * p).sum(-1).max(dim=1, keepdim=True)[0].unsqueeze(1)
    eps = tol * scale + 1e-30  # B,M,1
    idx = idx.long().cpu(); ref_idx = ref_idx.long().cpu(); dist = dist.cpu()
    assert idx.shape == ref_idx.shape, (idx.shape, ref_idx.shape)
    assert int(idx.min()) >= 0 and int(idx.max()) < points.size(2)
    # no index twice in a row of results
    srt = idx.sort(dim=-1)[0]
    assert bool((srt[..., 1:] != srt[..., :-1]).all()), "duplicate index inside one neighbourhood"
    d_mine = torch.gather(D, 2, idx)
    d_ref = torch.gather(D, 2, ref_idx)
    differ = idx != ref_idx
    bad = differ & ((d_mine - d_ref).abs() > eps)
    assert not bool(bad.any()), f"{int(bad.sum())} neighbour indices differ beyond the near-tie tolerance"
    # reported distances: the float64 value within rounding, and ascending
    assert bool(((dist.double() - d_mine).abs() <= eps).all()), "reported distances off"
    assert bool((dist[..., 1:] >= dist[..., :-1]).all()), "distances not ascending"
    return int(differ.sum())


def assert_close_frac(got, want, rtol=1e-5, atol=1e-6, frac=1.0, what=""):
    """|got-want| <= atol + rtol*|want| for at least `frac` of the elements (frac < 1 only where a discrete
    choice upstream -- a kNN near-tie -- may legitimately differ; the caller says why)."""
    got = torch.as_tensor(got).detach().cpu().double()
    want = torch.as_tensor(want).detach().cpu().double()
    assert got.shape == want.shape, (got.shape, want.shape)
    ok = (got - want).abs() <= atol + rtol * want.abs()
    share = ok.double().mean().item()
    worst = ((got - want).abs() / (atol + rtol * want.abs())).max().item()
    assert share >= frac, f"{what}: only {share:.6f} of elements within tolerance (need {frac}), worst ratio {worst:.3g}"
    return share


def cloud_match_fraction(a, b, tol):
    """a, b (3,N) point clouds: share of points of a that have a point of b within tol (and vice versa), min of both."""
    a = torch.as_tensor(a).double().t(); b = torch.as_tensor(b).double().t()
    d = torch.cdist(a, b)
    return min((d.min(dim=1)[0] <= tol).double().mean().item(), (d.min(dim=0)[0] <= tol).double().mean().item())
