"""GPU parity: the sm_100a kernels behind the reference's native entry points (sampling.*, losses.*)
against the CPU oracle (oracle/oracle_c.c).  Indices and distances bit-exact; atomically
accumulated gradients within 1e-5 relative (the reference's own order is unspecified)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests.util import bits_equal

pytestmark = pytest.mark.gpu


def _cloud(rng, b, n, dup_frac=0.0, scale=1.0):
    x = (rng.random((b, n, 3), dtype=np.float32) * 2 - 1) * scale
    if dup_frac > 0 and n > 1:
        cnt = max(1, int(n * dup_frac))
        for i in range(b):
            src = rng.integers(0, n, cnt)
            dst = rng.integers(0, n, cnt)
            x[i, dst] = x[i, src]
    return x


# ------------------------------------------------------------------------------------------- nmdistance
@pytest.mark.parametrize("b,n,m", [(2, 600, 700), (32, 624, 624), (1, 1, 1), (3, 513, 2049), (2, 5, 4100), (1, 3000, 7)])
def test_nmdistance_forward_bit_exact(pu3, cuda, b, n, m):
    rng = np.random.default_rng(n * 7 + m)
    x1, x2 = _cloud(rng, b, n, 0.05), _cloud(rng, b, m, 0.05)
    d1, i1, d2, i2 = c_oracle.nmdist_fwd(x1, x2)
    t1, t2 = torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda)
    o_d1 = torch.empty(b, n, device=cuda); o_i1 = torch.empty(b, n, dtype=torch.int32, device=cuda)
    o_d2 = torch.empty(b, m, device=cuda); o_i2 = torch.empty(b, m, dtype=torch.int32, device=cuda)
    assert pu3.losses.nmdistance_forward(t1, t2, o_d1, o_d2, o_i1, o_i2) == 1
    assert bits_equal(o_d1.cpu().numpy(), d1) and bits_equal(o_d2.cpu().numpy(), d2)
    assert np.array_equal(o_i1.cpu().numpy(), i1) and np.array_equal(o_i2.cpu().numpy(), i2)


def test_nmdistance_ties_pick_lowest_index(pu3, cuda):
    # every candidate identical: the reference's strict '<' keeps index 0 (nmdistance_cuda.cu:31,45,125)
    x1 = torch.zeros(1, 40, 3, device=cuda)
    x2 = torch.ones(1, 1100, 3, device=cuda)
    d1 = torch.empty(1, 40, device=cuda); i1 = torch.empty(1, 40, dtype=torch.int32, device=cuda)
    d2 = torch.empty(1, 1100, device=cuda); i2 = torch.empty(1, 1100, dtype=torch.int32, device=cuda)
    pu3.losses.nmdistance_forward(x1, x2, d1, d2, i1, i2)
    assert int(i1.abs().sum()) == 0 and int(i2.abs().sum()) == 0
    assert float(d1.min()) == 3.0 and float(d1.max()) == 3.0


@pytest.mark.parametrize("b,n,m", [(2, 600, 700), (32, 624, 624), (1, 1, 3)])
def test_nmdistance_backward(pu3, cuda, b, n, m):
    rng = np.random.default_rng(b + n + m)
    x1, x2 = _cloud(rng, b, n), _cloud(rng, b, m)
    g1 = rng.standard_normal((b, n)).astype(np.float32)
    g2 = rng.standard_normal((b, m)).astype(np.float32)
    _, i1, _, i2 = c_oracle.nmdist_fwd(x1, x2)
    e1, e2 = c_oracle.nmdist_bwd(x1, x2, g1, g2, i1, i2)
    tx1, tx2 = torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda)
    gx1, gx2 = torch.zeros_like(tx1), torch.zeros_like(tx2)
    pu3.losses.nmdistance_backward(tx1, tx2, gx1, gx2, torch.from_numpy(g1).to(cuda), torch.from_numpy(g2).to(cuda),
                                   torch.from_numpy(i1).to(cuda), torch.from_numpy(i2).to(cuda))
    np.testing.assert_allclose(gx1.cpu().numpy(), e1, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gx2.cpu().numpy(), e2, rtol=1e-5, atol=1e-6)


def test_nmdistance_backward_accepts_expanded_grads(pu3, cuda):
    # autograd of mean() hands over stride-0 gradients; the drop-in must not read them as dense garbage
    x1 = torch.rand(2, 50, 3, device=cuda); x2 = torch.rand(2, 60, 3, device=cuda)
    d1 = torch.empty(2, 50, device=cuda); i1 = torch.empty(2, 50, dtype=torch.int32, device=cuda)
    d2 = torch.empty(2, 60, device=cuda); i2 = torch.empty(2, 60, dtype=torch.int32, device=cuda)
    pu3.losses.nmdistance_forward(x1, x2, d1, d2, i1, i2)
    ga = torch.full((2, 1), 0.5, device=cuda).expand(2, 50)
    gb = torch.full((2, 1), 0.25, device=cuda).expand(2, 60)
    gx1, gx2 = torch.zeros_like(x1), torch.zeros_like(x2)
    pu3.losses.nmdistance_backward(x1, x2, gx1, gx2, ga, gb, i1, i2)
    e1, e2 = c_oracle.nmdist_bwd(x1.cpu().numpy(), x2.cpu().numpy(), ga.contiguous().cpu().numpy(),
                                 gb.contiguous().cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy())
    np.testing.assert_allclose(gx1.cpu().numpy(), e1, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gx2.cpu().numpy(), e2, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------- gather
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64])
@pytest.mark.parametrize("b,c,n,m", [(32, 3, 624, 1), (1, 3, 6240, 1248), (3, 24, 312, 500), (2, 1, 1, 4)])
def test_gather_forward_exact(pu3, cuda, dtype, b, c, n, m):
    g = torch.Generator().manual_seed(b * 1000 + n)
    pts = torch.randn(b, c, n, generator=g).to(dtype)
    idx = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32)
    want = c_oracle.gather_fwd(pts.numpy(), idx.numpy())
    out = torch.empty(b, c, m, dtype=dtype, device=cuda)
    ret = pu3.sampling.gather_forward(b, c, n, m, pts.to(cuda), idx.to(cuda), out)
    assert ret is out
    assert bits_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-12)])
def test_gather_backward(pu3, cuda, dtype, tol):
    b, c, n, m = 4, 5, 300, 900  # m > n: many collisions exercise the atomics
    g = torch.Generator().manual_seed(3)
    go = torch.randn(b, c, m, generator=g).to(dtype)
    idx = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32)
    want = c_oracle.gather_bwd(go.numpy(), idx.numpy(), n)
    gp = torch.zeros(b, c, n, dtype=dtype, device=cuda)
    pu3.sampling.gather_backward(b, c, n, m, go.to(cuda), idx.to(cuda), gp)
    np.testing.assert_allclose(gp.cpu().numpy(), want, rtol=tol, atol=tol)


def test_gather_points_autograd(pu3, cuda):
    feats = torch.randn(2, 3, 50, device=cuda, dtype=torch.float64, requires_grad=True)
    idx = torch.randint(0, 50, (2, 20), device=cuda, dtype=torch.int32)
    assert torch.autograd.gradcheck(pu3.operations.gather_points, (feats, idx), eps=1e-6, atol=1e-4)


# ------------------------------------------------------------------------------------------- FPS
def _run_fps(pu3, cuda, x, m, temp=None):
    b, n, _ = x.shape
    t = torch.from_numpy(x).to(cuda)
    idx = torch.empty(b, m, dtype=torch.int32, device=cuda)
    tt = torch.full((b, n), 1e10, device=cuda) if temp is None else torch.from_numpy(temp).to(cuda)
    ret = pu3.sampling.furthest_sampling(b, n, m, t, tt, idx)
    assert ret is idx
    return idx.cpu().numpy(), tt.cpu().numpy()


@pytest.mark.parametrize("b,n,m", [(1, 312, 64), (2, 624, 10), (1, 6240, 1248), (1, 1248, 20), (2, 5000, 48),
                                   (1, 1, 1), (1, 2, 2), (1, 3, 3), (1, 511, 100), (1, 512, 100), (1, 513, 100),
                                   (1, 8193, 300), (3, 40, 40), (1, 20, 16), (1, 12480, 2496)])
def test_fps_matches_oracle(pu3, cuda, b, n, m):
    rng = np.random.default_rng(n * 31 + m)
    x = _cloud(rng, b, n)
    want_temp = np.full((b, n), 1e10, np.float32)
    want = c_oracle.fps(x, m, temp=want_temp)
    got, got_temp = _run_fps(pu3, cuda, x, m)
    assert np.array_equal(got, want)
    assert bits_equal(got_temp, want_temp)  # temp is an in/out buffer in the reference ABI


@pytest.mark.parametrize("n,m,dup", [(700, 700, 0.3), (2000, 400, 0.5), (312, 312, 0.0)])
def test_fps_tie_rule_with_duplicates(pu3, cuda, n, m, dup):
    # duplicated points produce exact distance ties; the winner must follow the reference's
    # (k mod T, then k) rule (sampling_cuda.cu:147,162), and m == n forces an all-zero tail
    rng = np.random.default_rng(n)
    x = _cloud(rng, 2, n, dup_frac=dup)
    x = np.round(x * 8) / 8 if dup > 0 else x  # a coarse grid adds ties between distinct points too
    want = c_oracle.fps(x, m)
    got, _ = _run_fps(pu3, cuda, x, m)
    assert np.array_equal(got, want)


def test_fps_all_points_equal(pu3, cuda):
    x = np.ones((1, 600, 3), np.float32)
    want = c_oracle.fps(x, 50)
    got, _ = _run_fps(pu3, cuda, x, 50)
    assert np.array_equal(got, want)


def test_fps_batch_beyond_32_is_per_cloud(pu3, cuda):
    # the reference indexes temp by blockIdx.x and breaks for b > 32 (sampling_cuda.cu:131,146);
    # the replacement treats every cloud independently
    rng = np.random.default_rng(5)
    x = _cloud(rng, 40, 400)
    want = c_oracle.fps(x, 30, legacy_temp_rows=False)
    got, _ = _run_fps(pu3, cuda, x, 30)
    assert np.array_equal(got, want)
    for i in (0, 33, 39):
        assert np.array_equal(got[i], c_oracle.fps(x[i:i + 1], 30)[0])


def test_fps_honours_incoming_temp(pu3, cuda):
    rng = np.random.default_rng(9)
    x = _cloud(rng, 2, 900)
    temp = (rng.random((2, 900), dtype=np.float32) * 0.5).astype(np.float32)
    t_want = temp.copy()
    want = c_oracle.fps(x, 64, temp=t_want)
    got, t_got = _run_fps(pu3, cuda, x, 64, temp=temp.copy())
    assert np.array_equal(got, want) and bits_equal(t_got, t_want)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_fps_cluster_sizes_agree(pu3, cuda, cluster):
    rng = np.random.default_rng(77)
    x = _cloud(rng, 3, 6240, dup_frac=0.02)
    want = c_oracle.fps(x, 500)
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    lib.pu3_fps_set_cluster(cluster)
    try:
        got, _ = _run_fps(pu3, cuda, x, 500)
    finally:
        lib.pu3_fps_set_cluster(0)
    assert np.array_equal(got, want)


def test_fps_large_cloud_cluster_paths(pu3, cuda):
    rng = np.random.default_rng(123)
    x = _cloud(rng, 1, 24960)
    want = c_oracle.fps(x, 1200)
    got, _ = _run_fps(pu3, cuda, x, 1200)
    assert np.array_equal(got, want)
    # shared-memory-resident variant (cluster of 16, 32 points per thread)
    x = _cloud(rng, 1, 239616)
    want = c_oracle.fps(x, 120)
    got, _ = _run_fps(pu3, cuda, x, 120)
    assert np.array_equal(got, want)


def test_furthest_point_sample_wrapper(pu3, cuda):
    rng = np.random.default_rng(11)
    x = _cloud(rng, 2, 1000)
    want = c_oracle.fps(x, 100)
    xt = torch.from_numpy(x).to(cuda)
    idx, pts = pu3.operations.furthest_point_sample(xt.transpose(1, 2).contiguous(), 100, NCHW=True)
    assert idx.dtype == torch.int32 and np.array_equal(idx.cpu().numpy(), want)
    exp = np.take_along_axis(x, want[..., None].astype(np.int64), axis=1).transpose(0, 2, 1)
    assert bits_equal(pts.cpu().numpy(), np.ascontiguousarray(exp))
    idx2, pts2 = pu3.operations.furthest_point_sample(xt, 100, NCHW=False)
    assert np.array_equal(idx2.cpu().numpy(), want) and pts2.shape == (2, 100, 3)


# ------------------------------------------------------------------------------------------- ball query
def test_ball_query(pu3, cuda):
    rng = np.random.default_rng(2)
    xyz = _cloud(rng, 2, 300); q = _cloud(rng, 2, 40)
    want = c_oracle.ball_query(q, xyz, 0.4, 16)
    got = pu3.sampling.ball_query(torch.from_numpy(q).to(cuda), torch.from_numpy(xyz).to(cuda), 0.4, 16)
    assert np.array_equal(got.cpu().numpy(), want)


# ------------------------------------------------------------------------------------------- errors
def test_errors_raise_not_exit(pu3, cuda):
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pu3.sampling.furthest_sampling(1, 4, 2, torch.zeros(1, 4, 3), torch.zeros(1, 4), torch.zeros(1, 2, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="contiguous"):
        pu3.sampling.gather_forward(1, 3, 4, 2, torch.zeros(1, 4, 3, device=cuda).transpose(1, 2),
                                    torch.zeros(1, 2, dtype=torch.int32, device=cuda), torch.zeros(1, 3, 2, device=cuda))
    with pytest.raises(RuntimeError, match="empty cloud"):
        pu3._lib.check(pu3._lib.lib().pu3_fps_f32(1, 0, 3, None, None, None, None), "fps")


# ------------------------------------------------------------------------------------------- ragged batches
def test_fps_ragged_equals_per_cloud_calls(pu3, cuda):
    rng = np.random.default_rng(41)
    sizes = [700, 312, 6240, 513, 2000]
    wants = [64, 10, 300, 64, 1]
    nmax, mmax = max(sizes), max(wants)
    x = np.zeros((len(sizes), nmax, 3), np.float32)
    for i, n in enumerate(sizes):
        x[i, :n] = _cloud(rng, 1, n, dup_frac=0.05)[0]
        x[i, n:] = 1e6  # padding must never be looked at
    xt = torch.from_numpy(x).to(cuda).transpose(1, 2).contiguous()  # (B,3,Nmax)
    n_arr = torch.tensor(sizes, dtype=torch.int32, device=cuda)
    m_arr = torch.tensor(wants, dtype=torch.int32, device=cuda)
    idx, pts = pu3.operations.furthest_point_sample_ragged(xt, n_arr, m_arr, mmax)
    for i, (n, m) in enumerate(zip(sizes, wants)):
        want = c_oracle.fps(x[i:i + 1, :n], m)[0]
        assert np.array_equal(idx[i, :m].cpu().numpy(), want), i
        assert int(idx[i, m:].abs().sum()) == 0
        assert bits_equal(pts[i, :, :m].cpu().numpy(), np.ascontiguousarray(x[i, want].T))
    # more samples than points (a request whose tiles cover fewer points than the level must output)
    idx2, _ = pu3.operations.furthest_point_sample_ragged(xt, n_arr, None, 400)
    for i in (1, 3):
        assert np.array_equal(idx2[i].cpu().numpy(), c_oracle.fps(x[i:i + 1, :sizes[i]], 400)[0])


def test_furthest_point_sample_wrapper_beyond_on_chip_capacity(pu3, cuda):
    """clouds of more than 262144 points (the whole-shape resample of main.py:379 for shapes of >= 5.5k
    points) run from global memory and need `temp`; the operator allocates it like the reference (operations.py:291)."""
    n = pu3.operations.FPS_ONCHIP_MAX + 37
    rng = np.random.default_rng(4)
    x = _cloud(rng, 1, n)
    want = c_oracle.fps(x, 24)
    idx, pts = pu3.operations.furthest_point_sample(torch.from_numpy(x).to(cuda), 24, NCHW=False)
    assert np.array_equal(idx.cpu().numpy(), want)
    assert np.array_equal(pts.cpu().numpy()[0], x[0][want[0]])
