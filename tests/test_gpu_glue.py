"""GPU: the glue kernels of the eval path (csrc/glue.cu) against the torch expressions of the reference they replace
(network/operations.py:12-30, network/upsampler.py:63-85,138,144-158), evaluated on CPU by the oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_net

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,nchw", [((32, 3, 312), True), ((5, 3, 1), True), ((3, 3, 5000), True), ((4, 700, 3), False),
                                        ((1, 3, 24960), True)])
def test_normalize_point_batch_kernel(pu3, cuda, shape, nchw):
    """a-6: normalize_point_batch (operations.py:12-30) as one kernel, 1e-6 against the oracle's CPU evaluation."""
    g = torch.Generator().manual_seed(shape[0] * 7 + shape[-1])
    pc = torch.rand(*shape, generator=g) * 3 - 1
    want, wc, wr = ref_net.normalize_point_batch(pc, NCHW=nchw)
    got, gc, gr = pu3.operations.normalize_point_batch(pc.to(cuda), NCHW=nchw)
    assert got.shape == want.shape and gc.shape == wc.shape and gr.shape == wr.shape
    n = shape[2] if nchw else shape[1]
    if n == 1:     # a single point: 0 / 0 like the reference
        assert torch.isnan(got).all() and torch.isnan(want).all()
        return
    torch.testing.assert_close(gc.cpu(), wc, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(gr.cpu(), wr, rtol=1e-6, atol=0)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-6)
    assert abs(float(got.norm(dim=1 if nchw else 2).max()) - 1.0) < 1e-6


def test_normalize_point_batch_keeps_autograd(pu3, cuda):
    pc = torch.rand(2, 3, 50, device=cuda, requires_grad=True)
    out, c, r = pu3.operations.normalize_point_batch(pc)
    out.sum().backward()
    assert pc.grad is not None and bool(torch.isfinite(pc.grad).all())
    with torch.no_grad():
        k_out, k_c, k_r = pu3.operations.normalize_point_batch(pc.detach())     # kernel path
    torch.testing.assert_close(k_out, out.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(k_r, r.detach(), rtol=1e-6, atol=0)


def _outlier(pu3, cuda, xyz, d, k, r=2):
    B, _, N = xyz.shape
    dd = torch.stack([torch.zeros_like(d), d], dim=2).contiguous().to(cuda)
    cm, pm = torch.empty(B, 3, N, device=cuda), torch.empty(B, N, 3, device=cuda)
    arrs = [torch.empty(B, dtype=torch.int32, device=cuda) for _ in range(4)]
    bad = torch.zeros((), dtype=torch.int32, device=cuda)
    x = xyz.contiguous().to(cuda)
    pu3._lib.launch("pu3_outlier_compact_f32", x, B, N, 2, k, r, dd.data_ptr(), x.data_ptr(), cm.data_ptr(), pm.data_ptr(),
                    *[a.data_ptr() for a in arrs], bad.data_ptr())
    return cm.cpu(), pm.cpu(), [a.cpu() for a in arrs], int(bad)


@pytest.mark.parametrize("N", [624, 1248, 2496, 5000, 1000])
def test_outlier_filter_and_compaction(pu3, cuda, N):
    """upsampler.py:63-76: mask = d < 5 mean(d); masked_select keeps the order; patch_num = int(N'/k*5)."""
    g = torch.Generator().manual_seed(N)
    B, k = 3, 312
    xyz = torch.rand(B, 3, N, generator=g)
    d = torch.rand(B, N, generator=g) * 1e-3
    d[0, torch.randperm(N, generator=g)[:N // 9]] = 0.05          # removed points, scattered
    d[2, :7] = 1.0                                                # removed points at the front
    cm, pm, (n_arr, p_arr, pk_arr, pkr_arr), bad = _outlier(pu3, cuda, xyz, d, k)
    assert bad == 0
    for b in range(B):
        mask = d[b] < 5 * torch.mean(d[b])
        cnt = int(mask.sum())
        assert cnt < N or b == 1
        want = torch.masked_select(xyz[b], mask.unsqueeze(0).expand(3, -1)).view(3, -1)
        assert int(n_arr[b]) == cnt and int(p_arr[b]) == int(cnt / k * 5)
        assert int(pk_arr[b]) == int(p_arr[b]) * k and int(pkr_arr[b]) == int(p_arr[b]) * k * 2
        assert torch.equal(cm[b, :, :cnt], want) and torch.equal(pm[b, :cnt], want.t())
        rest = xyz[b][:, ~mask]                                   # removed points follow, order kept
        assert torch.equal(cm[b, :, cnt:], rest)


def test_outlier_filter_flags_a_cloud_smaller_than_one_tile(pu3, cuda):
    N, k = 624, 600
    xyz = torch.rand(1, 3, N)
    d = torch.ones(1, N); d[0, 100:200] = 100.0                   # 100 removed -> 524 kept < k
    _, _, (n_arr, p_arr, _, _), bad = _outlier(pu3, cuda, xyz, d, k)
    assert bad == 1 and int(n_arr[0]) == k and int(p_arr[0]) == int(524 / k * 5)


def test_tile_seeds_tiles_normalize_denorm_merge_gather(pu3, cuda):
    g = torch.Generator().manual_seed(5)
    B, N, P, k, r = 3, 700, 6, 64, 2
    L = pu3._lib
    xyz = torch.rand(B, 3, N, generator=g)
    idx = torch.randint(0, N, (B, P), generator=g, dtype=torch.int32)
    p_arr = torch.tensor([6, 4, 1], dtype=torch.int32)
    seeds = torch.empty(B, 3, P, device=cuda)
    x, idx_d, p_d = xyz.to(cuda), idx.to(cuda), p_arr.to(cuda)      # keep the device copies alive across the launch
    L.launch("pu3_tile_seeds_f32", x, B, N, P, x.data_ptr(), idx_d.data_ptr(), p_d.data_ptr(), seeds.data_ptr())
    for b in range(B):
        for j in range(P):
            src = int(idx[b, j if j < int(p_arr[b]) else 0])
            assert torch.equal(seeds[b, :, j].cpu(), xyz[b, :, src])
    # tiles (B,3,P,k) -> patches + normalisation + side-by-side cloud
    tiles = torch.rand(B, 3, P, k, generator=g) * 2 - 0.5
    t = tiles.to(cuda)
    T = B * P
    patch, pn = torch.empty(T, 3, k, device=cuda), torch.empty(T, 3, k, device=cuda)
    cen, rad = torch.empty(T, 3, device=cuda), torch.empty(T, device=cuda)
    sbs = torch.empty(B, 3, P * k, device=cuda)
    L.launch("pu3_tiles_normalize_f32", t, B, P, k, t.data_ptr(), patch.data_ptr(), pn.data_ptr(), cen.data_ptr(), rad.data_ptr(),
             sbs.data_ptr())
    want_patch = torch.cat(torch.unbind(tiles, dim=2), dim=0)      # upsampler.py:85 -- note: P-major ...
    want_patch = want_patch.view(P, B, 3, k).transpose(0, 1).reshape(T, 3, k)    # ... this library orders request-major
    assert torch.equal(patch.cpu(), want_patch)
    wn, wc, wr = ref_net.normalize_point_batch(want_patch)
    torch.testing.assert_close(pn.cpu(), wn, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(cen.cpu(), wc[:, :, 0], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(rad.cpu(), wr[:, 0, 0], rtol=1e-6, atol=0)
    assert torch.equal(sbs.cpu(), tiles.reshape(B, 3, P * k))
    # de-normalise + merge (point-major) and the gather that follows the merge FPS
    kr = k * r
    up = torch.rand(T, 3, kr, generator=g)
    merged = torch.empty(B, P * kr, 3, device=cuda)
    u = up.to(cuda)
    L.launch("pu3_denorm_merge_f32", u, B, P, kr, u.data_ptr(), cen.data_ptr(), rad.data_ptr(), merged.data_ptr())
    want = up * rad.cpu().view(T, 1, 1) + cen.cpu().view(T, 3, 1)                  # :144
    want = want.view(B, P, 3, kr).permute(0, 2, 1, 3).reshape(B, 3, P * kr)         # :149-155 per request
    assert torch.equal(merged.cpu().transpose(1, 2), want)
    m = 50
    oidx = torch.randint(0, P * kr, (B, m), generator=g, dtype=torch.int32)
    out = torch.empty(B, 3, m, device=cuda)
    oidx_d = oidx.to(cuda)
    L.launch("pu3_gather_pm_f32", merged, B, P * kr, m, merged.data_ptr(), oidx_d.data_ptr(), out.data_ptr())
    assert torch.equal(out.cpu(), torch.gather(want, 2, oidx.long().unsqueeze(1).expand(-1, 3, -1)))
