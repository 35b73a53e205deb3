"""GPU parity of group_knn (the reference runs it as PyTorch library calls + a CPU np.unique round trip,
network/operations.py:165-216) against the CPU oracle restatement oracle/ref_net.group_knn.

Indices: exact wherever the float64 distance gap between the two candidates exceeds the fp32 rounding of
the expanded-form distance (tests/util.knn_gap_check); neighbour features: bit-exact copies of
points[idx]; distances: within that same rounding, ascending."""
import numpy as np
import pytest
import torch

from oracle import ref_net
from tests.util import knn_gap_check, bits_equal

pytestmark = pytest.mark.gpu


def _check(pu3, cuda, k, q, p, unique, penalty_dups=False, max_group=None):
    pe = p if p.size(0) == q.size(0) else p.repeat_interleave(q.size(0) // p.size(0), dim=0)
    rk, ri, rd = ref_net.group_knn(k, q, pe, unique=unique, NCHW=True)
    nb, idx, dist = pu3.operations.group_knn(k, q.to(cuda), p.to(cuda), unique=unique, NCHW=True, max_group=max_group)
    assert nb.shape == rk.shape and idx.dtype == torch.int64 and idx.shape == ri.shape
    penalty = None
    if unique and penalty_dups:
        dmask = ref_net.duplicate_mask(pe.transpose(1, 2).contiguous()).double()
        D = ref_net.pairwise_sqdist_expanded(q.transpose(1, 2).contiguous(), pe.transpose(1, 2).contiguous())
        penalty = float(D.max()) * dmask
    ndiff = knn_gap_check(q, p, idx, dist, ri, k, penalty=penalty)
    # neighbour features are exact copies of the selected points
    B, C, M = q.shape
    exp = torch.gather(pe, 2, idx.cpu().view(B, 1, M * k).expand(-1, C, -1)).view(B, C, M, k)
    assert bits_equal(nb.cpu().numpy(), exp.numpy())
    return ndiff, idx.numel()


@pytest.mark.parametrize("b,c,n,k", [(4, 24, 312, 33), (2, 24, 312, 17), (3, 3, 700, 5), (2, 3, 2496, 2),
                                     (1, 3, 6240, 5), (2, 7, 100, 64), (2, 24, 40, 33), (1, 3, 33, 33)])
def test_self_knn_small_k(pu3, cuda, b, c, n, k):
    g = torch.Generator().manual_seed(b * 100 + n + k)
    x = torch.rand(b, c, n, generator=g)
    ndiff, total = _check(pu3, cuda, k, x, x, unique=True)
    assert ndiff <= max(4, total // 500), (ndiff, total)  # near-ties are rare on random data


@pytest.mark.parametrize("b,m,n,k", [(2, 5, 3000, 312), (2, 1, 4992, 2496), (3, 1, 624, 312), (1, 48, 5000, 312),
                                     (1, 2, 300, 300), (1, 3, 24960, 312), (2, 4, 1000, 65), (1, 2, 9000, 4992)])
def test_patch_extraction_large_k(pu3, cuda, b, m, n, k):
    g = torch.Generator().manual_seed(m * 10 + n + k)
    p = torch.rand(b, 3, n, generator=g)
    seeds = p[:, :, torch.randperm(n, generator=g)[:m]].contiguous()
    ndiff, total = _check(pu3, cuda, k, seeds, p, unique=False)
    assert ndiff <= max(4, total // 500), (ndiff, total)


def test_query_differs_from_points_with_shared_cloud(pu3, cuda):
    # inter-level skip connection: 10 patches of 312 queries against ONE previous-level cloud
    # (the reference expand()s it, upsampler.py:319-323); here the points batch divides the query batch
    g = torch.Generator().manual_seed(4)
    prev = torch.rand(2, 3, 3120, generator=g)
    q = torch.rand(20, 3, 312, generator=g)
    _check(pu3, cuda, 5, q, prev, unique=True)


def test_duplicates_are_pushed_back(pu3, cuda):
    g = torch.Generator().manual_seed(8)
    p = torch.rand(3, 3, 200, generator=g)
    p[0, :, 50] = p[0, :, 10]; p[0, :, 51] = p[0, :, 10]; p[1, :, 199] = p[1, :, 0]; p[2, :, 7] = p[2, :, 3]
    q = p[:, :, :64].contiguous()
    _check(pu3, cuda, 8, q, p, unique=True, penalty_dups=True)
    nb, idx, dist = pu3.operations.group_knn(200, q.to(cuda), p.to(cuda), unique=True)
    # with k == n every duplicate must sit at the very end of its row, after all first occurrences
    assert set(idx[0, 0, -2:].tolist()) == {50, 51} and int(idx[1, 0, -1]) == 199 and int(idx[2, 0, -1]) == 7
    # unique=False keeps them where their distance puts them: point 10's row has 10, 50, 51 in index order first
    _, idx2, d2 = pu3.operations.group_knn(3, q.to(cuda), p.to(cuda), unique=False)
    assert idx2[0, 10].tolist() == [10, 50, 51]


def test_duplicate_penalty_group_scope(pu3, cuda):
    # max(D) is taken over the whole batch in the reference (operations.py:204); max_group narrows it
    g = torch.Generator().manual_seed(9)
    p = torch.rand(4, 3, 100, generator=g)
    p[1] *= 10.0  # cloud 1 has by far the largest distances
    p[0, :, 5] = p[0, :, 2]
    pc = p.to(cuda)
    _, _, d_all = pu3.operations.group_knn(100, pc, pc, unique=True)
    _, _, d_own = pu3.operations.group_knn(100, pc, pc, unique=True, max_group=1)
    D = ref_net.pairwise_sqdist_expanded(p.transpose(1, 2).contiguous(), p.transpose(1, 2).contiguous())
    add_all = float(D.max()); add_own = float(D[0].max())
    assert abs(float(d_all[0, 0, -1]) - (float(D[0, 0, 5]) + add_all)) <= 1e-4 * add_all
    assert abs(float(d_own[0, 0, -1]) - (float(D[0, 0, 5]) + add_own)) <= 1e-4 * add_own


def test_nhwc_layout_and_errors(pu3, cuda):
    g = torch.Generator().manual_seed(1)
    p = torch.rand(2, 150, 3, generator=g); q = torch.rand(2, 9, 3, generator=g)
    rk, ri, rd = ref_net.group_knn(6, q, p, unique=True, NCHW=False)
    nb, idx, dist = pu3.operations.group_knn(6, q.to(cuda), p.to(cuda), unique=True, NCHW=False)
    assert nb.shape == (2, 9, 6, 3) and rk.shape == nb.shape
    knn_gap_check(q.transpose(1, 2).contiguous(), p.transpose(1, 2).contiguous(), idx, dist, ri, 6)
    exp = torch.gather(p.unsqueeze(1).expand(-1, 9, -1, -1), 2, idx.cpu().unsqueeze(-1).expand(-1, -1, -1, 3))
    assert bits_equal(nb.contiguous().cpu().numpy(), exp.contiguous().numpy())
    with pytest.raises(AssertionError, match="greater or equal to k"):
        pu3.operations.group_knn(200, q.to(cuda), p.to(cuda), NCHW=False)
    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        pu3.operations.group_knn(3, q, p, NCHW=False)


def test_group_knn_backward_matches_oracle_autograd(pu3, cuda):
    g = torch.Generator().manual_seed(12)
    p0 = torch.rand(2, 6, 80, generator=g); q0 = torch.rand(2, 6, 30, generator=g)
    w_nb = torch.randn(2, 6, 30, 7, generator=g); w_d = torch.randn(2, 30, 7, generator=g)

    def run(fn, dev):
        p = p0.clone().to(dev).requires_grad_(); q = q0.clone().to(dev).requires_grad_()
        nb, idx, d = fn(7, q, p, unique=True, NCHW=True)
        ((nb * w_nb.to(dev)).sum() + (d * w_d.to(dev)).sum()).backward()
        return p.grad.cpu(), q.grad.cpu(), idx.cpu()

    gp_ref, gq_ref, i_ref = run(ref_net.group_knn, "cpu")
    gp, gq, i = run(pu3.operations.group_knn, cuda)
    assert torch.equal(i, i_ref)
    torch.testing.assert_close(gp, gp_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gq, gq_ref, rtol=1e-5, atol=1e-5)


def test_chamfer_loss_module(pu3, cuda):
    g = torch.Generator().manual_seed(3)
    a0 = torch.rand(4, 3, 624, generator=g); b0 = torch.rand(4, 3, 624, generator=g)
    for thr in (None, 2.0):
        a = a0.clone().requires_grad_()
        want = ref_net.chamfer_loss(a, b0, threshold=thr)
        want.backward()
        ac = a0.clone().to(cuda).requires_grad_()
        got = pu3.model_loss.ChamferLoss(threshold=thr)(ac, b0.to(cuda))
        got.backward()
        torch.testing.assert_close(got.cpu(), want, rtol=1e-6, atol=1e-8)
        torch.testing.assert_close(ac.grad.cpu(), a.grad, rtol=1e-5, atol=1e-8)


def test_ragged_batch_equals_per_element_calls(pu3, cuda):
    """Batched eval: tiles of different requests in one launch, every request with its own cloud size."""
    g = torch.Generator().manual_seed(21)
    sizes = [900, 400, 1300]
    nmax = max(sizes)
    clouds = torch.full((3, 3, nmax), 50.0)
    for i, n in enumerate(sizes):
        clouds[i, :, :n] = torch.rand(3, n, generator=g)
    owner = torch.tensor([0, 0, 1, 2, 2, 2, 2], dtype=torch.int32)
    q = torch.rand(7, 3, 100, generator=g)
    R = pu3.operations.Ragged(owner.to(cuda), owner.to(cuda), 3, n_arr=torch.tensor(sizes, dtype=torch.int32, device=cuda))
    for k in (5, 200):
        nb, idx, dist = pu3.operations._knn_raw(k, q.to(cuda), clouds.to(cuda), True, None, ragged=R)
        for i in range(7):
            c = int(owner[i]); n = sizes[c]
            nb1, idx1, dist1 = pu3.operations.group_knn(k, q[i:i + 1].to(cuda), clouds[c:c + 1, :, :n].contiguous().to(cuda),
                                                        unique=True)
            assert torch.equal(idx[i:i + 1], idx1) and torch.equal(nb[i:i + 1], nb1) and torch.equal(dist[i:i + 1], dist1)
    # ragged query counts (seeds per request) with k > 64
    m_arr = torch.tensor([3, 1, 2], dtype=torch.int32, device=cuda)
    req = torch.arange(3, dtype=torch.int32, device=cuda)
    seeds = clouds[:, :, :3].contiguous()
    R2 = pu3.operations.Ragged(req, req, 3, n_arr=torch.tensor(sizes, dtype=torch.int32, device=cuda), m_arr=m_arr)
    nb, idx, _ = pu3.operations._knn_raw(312, seeds.to(cuda), clouds.to(cuda), False, None, ragged=R2)
    for i in range(3):
        n = sizes[i]; m = int(m_arr[i])
        nb1, idx1, _ = pu3.operations.group_knn(312, seeds[i:i + 1, :, :m].contiguous().to(cuda),
                                                clouds[i:i + 1, :, :n].contiguous().to(cuda), unique=False)
        assert torch.equal(idx[i:i + 1, :m], idx1) and torch.equal(nb[i:i + 1, :, :m], nb1)
        assert int(idx[i, m:].abs().sum()) == 0


def test_tiled_and_streaming_kernels_agree(pu3, cuda):
    """k <= 64 has two kernels (tiled sort+pop for clouds of <= 320 points, streaming insertion otherwise):
    identical results, including tie order and the duplicate penalty."""
    import ctypes
    g = torch.Generator().manual_seed(33)
    x = torch.rand(5, 24, 312, generator=g)
    x[0, :, 100] = x[0, :, 7]; x[3, :, 311] = x[3, :, 0]
    q = torch.rand(5, 24, 40, generator=g)
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    outs = []
    for force in (0, 1):
        lib.pu3_knn_force_stream(force)
        try:
            outs.append((pu3.operations.group_knn(33, x.to(cuda), x.to(cuda), unique=True),
                         pu3.operations.group_knn(9, q.to(cuda), x.to(cuda), unique=True),
                         pu3.operations.group_knn(64, x[:, :3, :70].contiguous().to(cuda), x[:, :3, :70].contiguous().to(cuda), unique=False),
                         # thread-per-query kernel (3 channels, k <= 8) against the streaming kernel
                         pu3.operations.group_knn(5, q[:, :3].contiguous().to(cuda), x[:, :3].contiguous().to(cuda), unique=True),
                         pu3.operations.group_knn(2, x[:, :3].contiguous().to(cuda), x[:, :3].contiguous().to(cuda), unique=False),
                         pu3.operations.group_knn(8, x[:, :3].contiguous().to(cuda), x[:, :3].contiguous().to(cuda), unique=True)))
        finally:
            lib.pu3_knn_force_stream(0)
    for a, b in zip(outs[0], outs[1]):
        for ta, tb in zip(a, b):
            assert torch.equal(ta, tb)


@pytest.mark.parametrize("n", [40, 312, 1025, 3120, 6240, 16384])
def test_large_cloud_duplicate_detection_hash_equals_scan_and_oracle(pu3, cuda, n):
    """Clouds of up to 16384 points mark duplicates with a shared-memory hash table; the O(n^2) scans (test hook) and the
    oracle (np.unique first occurrences, operations.py:192-204) must give the same neighbours, also with many
    duplicates, -0.0 / +0.0 and fewer than k distinct points (exact max(D) penalty path)."""
    import ctypes
    g = torch.Generator().manual_seed(n)
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    pts = torch.rand(2, 3, n, generator=g)
    src = torch.randint(0, n // 2, (n // 3,), generator=g)
    pts[0, :, n - n // 3:] = pts[0, :, src]            # a third of cloud 0 are copies of earlier points
    pts[0, 0, 5] = 0.0; pts[0, :, 9] = pts[0, :, 5]; pts[0, 0, 9] = -0.0   # equal as values, different bits
    pts[1, :, 4:] = pts[1, :, :4].repeat(1, (n + 3) // 4)[:, :n - 4]     # cloud 1: only 4 distinct points < k
    q = torch.rand(2, 3, 50, generator=g)
    outs = []
    for scan in (0, 1):
        lib.pu3_knn_dup_scan(scan)
        try:
            outs.append(pu3.operations.group_knn(5, q.to(cuda), pts.to(cuda), unique=True, max_group=1))
        finally:
            lib.pu3_knn_dup_scan(0)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for i in range(2):   # max_group=1: every cloud is its own request, like one reference call per cloud
        _, ridx, rd = ref_net.group_knn(5, q[i:i + 1], pts[i:i + 1], unique=True)
        got_idx, got_d = outs[0][1][i:i + 1].cpu(), outs[0][2][i:i + 1].cpu()
        cols = slice(0, 5) if i == 0 else slice(0, 4)      # cloud 1: rank 4 is a tie between ~n/4 copies (topk order unspecified)
        same = (got_idx[..., cols] == ridx[..., cols]).float().mean().item()
        assert same > 0.99, same                                            # near-ties aside, the same neighbours
        assert torch.allclose(got_d, rd, rtol=1e-4, atol=1e-6)             # incl. the + max(D) of the penalised ranks


def test_set_order_mode_returns_the_same_neighbour_sets(pu3, cuda):
    """PU3_KNN_SET_ORDER (the feature kNN of the level engine): rank 0 identical to the exact mode, ranks 1..k-1 the
    same set; clouds with duplicates and with massive ties (all points equal, few distinct points) take the exact
    fallback inside the kernel and must agree as well."""
    g = torch.Generator().manual_seed(44)
    x = torch.rand(6, 24, 312, generator=g)
    x[1, :, 200:] = x[1, :, :112]                       # duplicates: pushed behind the first occurrences
    x[2] = x[2, :, :1].expand(-1, 312)                  # all points equal: every distance ties
    x[3, :, 5:] = x[3, :, :5].repeat(1, 62)[:, :307]    # 5 distinct points < k
    x[4] = torch.relu(x[4] - 0.7)                       # post-ReLU features: many exact zeros, clamped distances
    xc = x.to(cuda)
    for k in (33, 17, 2):
        _, exact, _ = pu3.operations._knn_raw(k, xc, xc, True, 1, want_knn=False, want_dist=False, idx_dtype=torch.int32)
        _, fast, _ = pu3.operations._knn_raw(k, xc, xc, True, 1, want_knn=False, want_dist=False, idx_dtype=torch.int32,
                                             set_order=True)
        assert torch.equal(exact[..., 0], fast[..., 0])                               # rank 0
        assert torch.equal(exact.sort(dim=-1)[0], fast.sort(dim=-1)[0]), k             # the same k neighbours
    # the edge-conv built on either index list is bit-identical (max over the neighbours is order-free)
    params = ref_net.make_params(1, seed=1)
    pre = "levels.level_1.layer1"
    ws = [params[f"{pre}.mlps.{i}.weight"].to(cuda) for i in range(3)]
    bs = [params[f"{pre}.mlps.{i}.bias"].to(cuda) for i in range(3)]
    _, exact, _ = pu3.operations._knn_raw(33, xc, xc, True, 1, want_knn=False, want_dist=False, idx_dtype=torch.int32)
    _, fast, _ = pu3.operations._knn_raw(33, xc, xc, True, 1, want_knn=False, want_dist=False, idx_dtype=torch.int32, set_order=True)
    with torch.no_grad():
        a = pu3.fused.dense_edge_conv(xc, ws, bs, 32, idx=exact[..., 1:].long())[0]
        b = pu3.fused.dense_edge_conv(xc, ws, bs, 32, idx=fast[..., 1:].long())[0]
    assert torch.equal(a, b)


@pytest.mark.parametrize("case", ["tile_in_big_cloud", "queries_spread_out", "cloud_far_away", "ragged"])
def test_knn_thread_prefilter_is_exact(pu3, cuda, case):
    """The skip connection's search (c=3, k<=8, thousands of candidates) pre-filters candidates with a bounding sphere around the
    CTA's queries and verifies the result afterwards (csrc/group_knn.cu); whatever the geometry -- including the ones where the
    verification must fail and the CTA redoes the search unfiltered -- indices and distances are those of the unfiltered kernel,
    bit for bit."""
    import ctypes
    g = torch.Generator().manual_seed(hash(case) % 1000)
    B, M, N, k = 6, 312, 6240, 5
    base = torch.rand(B, 3, N // 5, generator=g)
    pts = base.repeat(1, 1, 5)                                     # 5x duplicated, like the merged previous-level tiles
    if case == "tile_in_big_cloud":
        centre = base[:, :, :1]
        d = ((base - centre) ** 2).sum(1)
        near = d.argsort(dim=1)[:, :M]                              # the M points nearest to one of the cloud's points: a tile
        q = torch.gather(base, 2, near.unsqueeze(1).expand(-1, 3, -1)) + 1e-3 * torch.randn(B, 3, M, generator=g)
    elif case == "queries_spread_out":
        q = torch.rand(B, 3, M, generator=g)
    elif case == "cloud_far_away":
        q = torch.rand(B, 3, M, generator=g) * 0.05 + 5.0            # nothing inside the sphere: every CTA must fall back
    else:
        q = torch.rand(B, 3, M, generator=g) * 0.2 + 0.4
    q, pts = q.contiguous().to(cuda), pts.contiguous().to(cuda)
    ragged = None
    if case == "ragged":
        owner = torch.tensor([0, 0, 1, 1, 2, 2], dtype=torch.int32, device=cuda)
        n_arr = torch.tensor([6240, 3120, 1248], dtype=torch.int32, device=cuda)
        ragged = pu3.operations.Ragged(owner, owner, 3, n_arr=n_arr)
        pts = pts[:3].contiguous()
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    outs = []
    for off in (0, 1):
        lib.pu3_knn_no_prefilter(off)
        try:
            _, idx, dist = pu3.operations._knn_raw(k, q, pts, True, 1, want_knn=False, ragged=ragged)
            torch.cuda.synchronize()
        finally:
            lib.pu3_knn_no_prefilter(0)
        outs.append((idx.clone(), dist.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert int(outs[0][0].max()) < N


@pytest.mark.parametrize("case", ["uniform", "tiles_in_merged_cloud", "queries_outside", "clustered", "surface", "ragged", "self_k2",
                                  "degenerate", "k8_with_dist"])
def test_knn_grid_search_is_exact(pu3, cuda, case):
    """xyz searches over >= 1024 candidates (skip connection at levels 3 / 4, outlier filter of the merged clouds) sort the
    candidates into a uniform grid and look at the 27 cells around a query, verify the k-th distance against the distance to the
    unscanned cells, widen to 5x5x5 and fall back to all cells otherwise (csrc/group_knn.cu: knn_grid_kernel).  Whatever the
    geometry -- uniform volumes, 5x duplicated merged clouds, queries outside the candidates' box (every query falls back), one
    dense cluster plus far outliers (most points in one cell), a thin surface, ragged clouds, the self search of the outlier
    filter, a degenerate cloud with fewer than k distinct points (exact max(D) penalty) -- indices, distances and gathered
    neighbours are those of the exhaustive kernel, bit for bit."""
    import ctypes
    g = torch.Generator().manual_seed(sum(map(ord, case)))
    B, M, N, k = 4, 312, 6240, 5
    unique, ragged, want_knn = True, None, False
    if case == "uniform":
        pts = torch.rand(B, 3, N, generator=g); q = torch.rand(B, 3, M, generator=g)
    elif case == "tiles_in_merged_cloud":
        base = torch.rand(B, 3, N // 5, generator=g)
        pts = base.repeat(1, 1, 5)[:, :, torch.randperm(N, generator=g)]
        d = ((base - base[:, :, :1]) ** 2).sum(1)
        near = d.argsort(dim=1)[:, :M]
        q = torch.gather(base, 2, near.unsqueeze(1).expand(-1, 3, -1)) + 1e-3 * torch.randn(B, 3, M, generator=g)
    elif case == "queries_outside":
        pts = torch.rand(B, 3, N, generator=g); q = torch.rand(B, 3, M, generator=g) * 0.1 + 3.0
    elif case == "clustered":
        pts = torch.cat([torch.randn(B, 3, N - 40, generator=g) * 0.01, torch.rand(B, 3, 40, generator=g) * 10 - 5], 2)
        q = torch.cat([torch.randn(B, 3, M - 12, generator=g) * 0.01, torch.rand(B, 3, 12, generator=g) * 10 - 5], 2)
    elif case == "surface":
        uv = torch.rand(B, 2, N, generator=g)
        pts = torch.stack([uv[:, 0], uv[:, 1], 0.2 * torch.sin(6 * uv[:, 0]) * torch.cos(5 * uv[:, 1])], 1)
        q = pts[:, :, :M] + 2e-3 * torch.randn(B, 3, M, generator=g)
    elif case == "ragged":
        pts = torch.rand(3, 3, N, generator=g); q = torch.rand(6, 3, M, generator=g)
        owner = torch.tensor([0, 0, 1, 2, 2, 2], dtype=torch.int32, device=cuda)
        ragged = pu3.operations.Ragged(owner, owner, 3, n_arr=torch.tensor([6240, 3000, 1100], dtype=torch.int32, device=cuda))
    elif case == "self_k2":
        N, k = 2496, 2
        pts = torch.rand(B, 3, N, generator=g); pts[:, :, 100:110] = pts[:, :, :10]          # a few exact duplicates
        q = None
    elif case == "degenerate":
        N = 1024
        pts = torch.rand(B, 3, 3, generator=g).repeat(1, 1, 342)[:, :, :N].contiguous()      # 3 distinct points < k
        q = torch.rand(B, 3, M, generator=g)
    else:
        k, want_knn, unique = 8, True, False
        pts = torch.rand(B, 3, N, generator=g); q = torch.rand(B, 3, 700, generator=g)
    pts = pts.contiguous().to(cuda)
    q = pts if q is None else q.contiguous().to(cuda)
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    outs = []
    for on in (1, 0):
        lib.pu3_knn_set_grid(on)
        try:
            nb, idx, dist = pu3.operations._knn_raw(k, q, pts, unique, 1, want_knn=want_knn, ragged=ragged)
            torch.cuda.synchronize()
        finally:
            lib.pu3_knn_set_grid(1)
        outs.append((idx.clone(), dist.clone(), None if nb is None else nb.clone()))
    assert torch.equal(outs[0][0], outs[1][0]), f"{case}: {(outs[0][0] != outs[1][0]).sum().item()} indices differ"
    assert torch.equal(outs[0][1], outs[1][1])
    if want_knn:
        assert torch.equal(outs[0][2], outs[1][2])
    assert int(outs[0][0].max()) < pts.shape[2]


@pytest.mark.parametrize("case", ["no_dups", "some_dups", "degenerate", "all_equal", "ragged_groups"])
def test_feature_knn_fused_duplicate_detection_equals_the_prepass(pu3, cuda, case):
    """The tiled feature-space kernel finds duplicates itself (in-kernel hash compare; exact max(D) penalty computed in the
    kernel for a degenerate cloud); results must be those of the three-kernel pre-pass (hash table, group flags, max D) it
    replaces -- indices, distances and gathered neighbours, bit for bit -- and the oracle's."""
    import ctypes
    g = torch.Generator().manual_seed(len(case))
    B, C, N, k = 6, 24, 312, 33
    x = torch.relu(torch.randn(B, C, N, generator=g))            # post-ReLU features: many exact zeros
    if case == "some_dups":
        x[:, :, 100:140] = x[:, :, 0:40]                          # 40 duplicated points per cloud
        x[2, :, 200:260] = x[2, :, 5:6]                           # and a 60-fold copy of one point
    elif case == "degenerate":
        x[1, :, 20:] = x[1, :, 3:4]                               # cloud 1: 21 distinct points < k -> exact max(D) penalty
        x[4, :, 10:] = x[4, :, :10].repeat(1, 31)[:, :N - 10]     # cloud 4: 10 distinct points
    elif case == "all_equal":
        x[:] = 0.5
    x = x.contiguous().to(cuda)
    ragged = None
    max_group = 2 if case != "ragged_groups" else None
    if case == "ragged_groups":
        x[0, :, 150:] = x[0, :, 7:8]                              # a degenerate cloud inside a three-cloud group
        me = torch.arange(B, dtype=torch.int32, device=cuda)
        grp = torch.tensor([0, 0, 0, 1, 1, 1], dtype=torch.int32, device=cuda)
        ragged = pu3.operations.Ragged(me, grp, 2)
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    outs = []
    for off in (0, 1):
        lib.pu3_knn_no_fused_dup(off)
        try:
            knn, idx, dist = pu3.operations._knn_raw(k, x, x, True, max_group, ragged=ragged)
            _, idx32, _ = pu3.operations._knn_raw(k, x, x, True, max_group, want_knn=False, want_dist=False, idx_dtype=torch.int32,
                                                  ragged=ragged, set_order=True)
            torch.cuda.synchronize()
        finally:
            lib.pu3_knn_no_fused_dup(0)
        outs.append((knn.clone(), idx.clone(), dist.clone(), idx32.sort(dim=2)[0]))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    if case in ("no_dups", "some_dups") :                         # and against the oracle (whole-batch penalty scope = one group)
        _, ridx, rdist = ref_net.group_knn(k, x.cpu()[:2], x.cpu()[:2], unique=True)
        _, gidx, gdist = pu3.operations._knn_raw(k, x[:2].contiguous(), x[:2].contiguous(), True, 2)
        assert (gidx.cpu() == ridx).float().mean() > 0.999
        torch.testing.assert_close(gdist.cpu(), rdist, rtol=1e-4, atol=5e-5)   # rank 0 is the self-distance: rounding noise of |x|^2 ~ 12
