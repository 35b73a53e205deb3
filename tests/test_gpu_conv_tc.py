"""GPU parity of the tensor-core (tcgen05, 3xTF32) 1x1 convolutions of the expansion head (csrc/conv_tc.cu)
against a float64 CPU evaluation of the reference's operators (nn.Conv2d 1x1 + ReLU, network/layers.py:161-204;
feature expansion and head, network/upsampler.py:349-372).

Tolerance: 1e-5 relative (+1e-5 absolute near zero) -- the bar BASELINE.json states for float features, the same one
the FFMA kernels are held to (tests/test_gpu_network.py::test_pointwise_conv).  A plain TF32 product would miss it by
two orders of magnitude; the hi/lo split is what is being tested."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_net
from tests.util import assert_close_frac

pytestmark = pytest.mark.gpu


def _rand(g, *shape, scale=1.0):
    return torch.randn(*shape, generator=g) * scale


# ragged point counts (tails of the 128-point MMA tile, clouds smaller than a TMA box), channel counts that are
# not a multiple of the 32-channel stage or of the 8-channel MMA, cout that is not a multiple of 16
@pytest.mark.parametrize("b,n,cin,cout,relu", [(3, 624, 128, 128, True), (2, 312, 264, 128, False), (5, 312, 204, 24, True),
                                               (1, 40, 84, 24, True), (7, 100, 8, 64, False), (2, 4, 3, 1, False),
                                               (3, 128, 33, 128, True), (2, 132, 129, 65, True), (301, 624, 128, 64, True)])
def test_conv_tc_plain(pu3, cuda, b, n, cin, cout, relu):
    g = torch.Generator().manual_seed(cin * cout + n)
    x, w, bias = _rand(g, b, cin, n), _rand(g, cout, cin, scale=0.2), _rand(g, cout)
    want = F.conv1d(x.double(), w.double().unsqueeze(-1), bias.double())
    want = F.relu(want) if relu else want
    out = torch.full((b, cout, n), float("nan"), device=cuda)
    pu3.fused.tc_conv_into(x.to(cuda), w.to(cuda), bias.to(cuda), out, relu=relu)
    assert_close_frac(out, want, rtol=1e-5, atol=1e-5, what="conv_tc")


def test_conv_tc_matches_ffma_kernel_closely(pu3, cuda):
    """Both kernels round differently (different summation order, split products) but agree far inside the tolerance."""
    g = torch.Generator().manual_seed(11)
    x, w, bias = _rand(g, 16, 128, 624).to(cuda), _rand(g, 128, 128, scale=0.1).to(cuda), _rand(g, 128).to(cuda)
    a = torch.empty(16, 128, 624, device=cuda); b = torch.empty_like(a)
    pu3.fused.tc_conv_into(x, w, bias, a, relu=True)
    pu3.fused.conv_into(x, w, bias, b, relu=True)
    assert float((a - b).abs().max()) < 2e-5 * float(b.abs().max())


def test_conv_tc_channel_slices_and_no_bias(pu3, cuda):
    g = torch.Generator().manual_seed(5)
    buf = _rand(g, 4, 264, 312).to(cuda)
    out = torch.zeros(4, 50, 312, device=cuda)
    w = _rand(g, 24, 204, scale=0.1).to(cuda)
    pu3.fused.tc_conv_into(buf[:, 60:], w, None, out[:, 10:34], relu=True)
    want = F.relu(F.conv1d(buf[:, 60:].cpu().double(), w.cpu().double().unsqueeze(-1)))
    assert_close_frac(out[:, 10:34], want, atol=1e-5)
    assert float(out[:, :10].abs().sum()) == 0 and float(out[:, 34:].abs().sum()) == 0   # neighbours untouched


@pytest.mark.parametrize("b,n,cin,cout,r", [(3, 312, 264, 128, 2), (2, 100, 40, 64, 3), (1, 8, 264, 128, 2), (2, 52, 7, 9, 4)])
def test_conv_tc_expand(pu3, cuda, b, n, cin, cout, r):
    """Feature expansion (upsampler.py:349-366): replicate every point r times, append the 1-D code, 1x1 conv, ReLU."""
    g = torch.Generator().manual_seed(n + r)
    x, w, bias = _rand(g, b, cin, n), _rand(g, cout, cin + 1, scale=0.2), _rand(g, cout)
    code = torch.linspace(-0.2, 0.2, r)
    rep = x.double().unsqueeze(-1).expand(-1, -1, -1, r).reshape(b, cin, n * r)                 # :352-353
    full = torch.cat([rep, code.double().repeat(n).view(1, 1, n * r).expand(b, -1, -1)], dim=1)  # :354-362
    want = F.relu(F.conv1d(full, w.double().unsqueeze(-1), bias.double()))
    got = pu3.fused.tc_expand(x.to(cuda), w.to(cuda), bias.to(cuda), code.to(cuda), r)
    assert got.shape == (b, cout, n * r)
    assert_close_frac(got, want, rtol=1e-5, atol=1e-5, what="conv_tc_expand")


@pytest.mark.parametrize("b,n,cin,cmid,cout,div", [(3, 624, 128, 64, 3, 2), (2, 52, 16, 40, 2, 1), (1, 4, 128, 64, 3, 4)])
def test_conv_tc_project(pu3, cuda, b, n, cin, cmid, cout, div):
    """fc_layer1 + ReLU + fc_layer2 + residual (upsampler.py:369-372) in one kernel."""
    g = torch.Generator().manual_seed(n + cmid)
    x, wm, bm = _rand(g, b, cin, n), _rand(g, cmid, cin, scale=0.2), _rand(g, cmid)
    wo, bo, res = _rand(g, cout, cmid, scale=0.2), _rand(g, cout), _rand(g, b, cout, n // div)
    h = F.relu(F.conv1d(x.double(), wm.double().unsqueeze(-1), bm.double()))
    want = F.conv1d(h, wo.double().unsqueeze(-1), bo.double()) + res.double().repeat_interleave(div, dim=2)
    got = pu3.fused.tc_project(x.to(cuda), wm.to(cuda), bm.to(cuda), wo.to(cuda), bo.to(cuda), residual=res.to(cuda), res_div=div)
    assert_close_frac(got, want, rtol=1e-5, atol=1e-5, what="conv_tc_project")
    got2 = pu3.fused.tc_project(x.to(cuda), wm.to(cuda), bm.to(cuda), wo.to(cuda), bo.to(cuda))
    assert_close_frac(got2, want - res.double().repeat_interleave(div, dim=2), rtol=1e-5, atol=1e-5, what="no residual")


def test_conv_tc_rejects_what_tma_cannot_address(pu3, cuda):
    x = torch.zeros(2, 8, 30, device=cuda)      # 30 points: rows are not 16-byte multiples
    with pytest.raises(RuntimeError, match="TMA"):
        pu3.fused.tc_conv_into(x, torch.zeros(8, 8, device=cuda), None, torch.empty(2, 8, 30, device=cuda))
    with pytest.raises(RuntimeError):
        pu3.fused.tc_prepare(torch.zeros(130, 8, device=cuda))   # cout > 128


def _head_reference(x, w1, b1, code, w2, b2, w3, b3, w4, b4, res):
    """float64 evaluation of upsampler.py:349-372 for step ratio 2: replicate, append the code, four 1x1 convolutions, residual."""
    B, C, N = x.shape
    xr = x.double().repeat_interleave(2, dim=2)                                  # (B,C,2N): point p -> 2p, 2p+1
    cd = code.double().repeat(N).view(1, 1, 2 * N).expand(B, 1, 2 * N)
    h = F.relu(F.conv1d(torch.cat([xr, cd], 1), w1.double().unsqueeze(-1), b1.double()))
    h = F.relu(F.conv1d(h, w2.double().unsqueeze(-1), b2.double()))
    h = F.relu(F.conv1d(h, w3.double().unsqueeze(-1), b3.double()))
    y = F.conv1d(h, w4.double().unsqueeze(-1), b4.double())
    return y + res.double().repeat_interleave(2, dim=2) if res is not None else y


# one tile on one CTA; ragged clouds (box tails, clouds smaller than a TMA box, boxes of different clouds in one tile); cin that
# is not a multiple of the 32-channel k-block; more tiles than SMs (every CTA runs several tiles: TMEM regions swap roles, the
# operand ring changes hands between the converters and the epilogue groups)
@pytest.mark.parametrize("b,n,cin,res", [(1, 128, 264, True), (3, 312, 264, True), (5, 100, 264, False), (7, 36, 72, True),
                                         (2, 4, 264, True), (40, 312, 264, True), (700, 312, 264, True)])
@pytest.mark.parametrize("mode", [2, 1, 0])
def test_head_tc_fused_chain(pu3, cuda, b, n, cin, res, mode):
    """mode 2 (default): CTA pairs (tcgen05.mma.cta_group::2), operands in tensor memory; mode 1: single CTAs, operands in tensor
    memory (TS form); mode 0: operands in shared memory."""
    pu3._lib.lib().pu3_head_tc_set_mode(mode)
    try:
        _head_case(pu3, cuda, b, n, cin, res)
    finally:
        pu3._lib.lib().pu3_head_tc_set_mode(2)


def _head_case(pu3, cuda, b, n, cin, res):
    g = torch.Generator().manual_seed(b * 1000 + n + cin)
    x = _rand(g, b, cin, n)
    w1, b1 = _rand(g, 128, cin + 1, scale=(2.0 / cin) ** 0.5), _rand(g, 128, scale=0.1)
    w2, b2 = _rand(g, 128, 128, scale=0.125), _rand(g, 128, scale=0.1)
    w3, b3 = _rand(g, 64, 128, scale=0.125), _rand(g, 64, scale=0.1)
    w4, b4 = _rand(g, 3, 64, scale=0.2), _rand(g, 3, scale=0.1)
    code = torch.tensor([-1.0, 1.0])
    r = _rand(g, b, 3, n) if res else None
    nref = min(b, 48)                                                            # float64 reference on a subset of the clouds
    pick = torch.linspace(0, b - 1, nref).round().long().unique()
    want = _head_reference(x[pick], w1, b1, code, w2, b2, w3, b3, w4, b4, None if r is None else r[pick])
    c = lambda t: None if t is None else t.to(cuda)
    out = pu3.fused.tc_head(c(x), c(w1), c(b1), c(code), c(w2), c(b2), c(w3), c(b3), c(w4), c(b4), residual=c(r))
    torch.cuda.synchronize()
    assert out.shape == (b, 3, 2 * n) and bool(torch.isfinite(out).all())
    assert_close_frac(out[pick.to(cuda)], want, rtol=1e-5, atol=1e-5, what="fused tcgen05 head")
    # and against the three-kernel tensor-core head (same 3xTF32 arithmetic up to the accumulator layout of up2)
    h1 = pu3.fused.tc_expand(c(x), c(w1), c(b1), c(code), 2)
    h2 = torch.empty_like(h1)
    pu3.fused.tc_conv_into(h1, c(w2), c(b2), h2, relu=True)
    three = pu3.fused.tc_project(h2, c(w3), c(b3), c(w4), c(b4), residual=c(r), res_div=2)
    assert float((out - three).abs().max()) <= 1e-5 * (1.0 + float(three.abs().max()))


def test_level_head_on_tensor_cores_matches_ffma_head_and_oracle(pu3, cuda):
    """The level engine with the tcgen05 head (fused and three-kernel) against (a) the same engine with the FFMA head and
    (b) the oracle."""
    params = ref_net.make_params(1, seed=1)
    net = pu3.Net(max_up_ratio=2, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
    net.load_state_dict(params, strict=True)
    net = net.to(cuda).eval()
    g = torch.Generator().manual_seed(7)
    xyz = ref_net.normalize_point_batch(torch.rand(5, 3, 312, generator=g))[0]
    lib = pu3._lib.lib()
    try:
        with torch.no_grad():
            lib.pu3_level_set_tc(3)          # default: fused head, prep convolutions on tensor cores
            all_xyz, all_feat = net.levels["level_1"](xyz.to(cuda), xyz.to(cuda))
            lib.pu3_level_set_tc(2)          # three-kernel head, prep convolutions on tensor cores
            k3_xyz, k3_feat = net.levels["level_1"](xyz.to(cuda), xyz.to(cuda))
            lib.pu3_level_set_tc(1)          # three-kernel head only
            tc_xyz, tc_feat = net.levels["level_1"](xyz.to(cuda), xyz.to(cuda))
            lib.pu3_level_set_tc(0)
            ff_xyz, ff_feat = net.levels["level_1"](xyz.to(cuda), xyz.to(cuda))
    finally:
        lib.pu3_level_set_tc(3)
    assert torch.equal(tc_feat, ff_feat)                                # the head does not touch the features
    assert torch.equal(all_feat, k3_feat)
    assert_close_frac(tc_xyz, ff_xyz, rtol=1e-5, atol=2e-6, what="tcgen05 head vs FFMA head")
    assert_close_frac(all_xyz, k3_xyz, rtol=1e-5, atol=2e-6, what="fused tcgen05 head vs three-kernel tcgen05 head")
    want_xyz, _ = ref_net.level_forward(params, "levels.level_1", xyz, xyz, None, knn=32)
    assert_close_frac(tc_xyz, want_xyz, rtol=1e-5, atol=2e-6, frac=0.99, what="tcgen05 head vs oracle")
    # prep convolutions on tensor cores feed the feature kNN: near-tie flips may move a small share of the features
    want_xyz2, want_feat = ref_net.level_forward(params, "levels.level_1", xyz, xyz, None, knn=32)
    assert_close_frac(all_feat, want_feat, rtol=1e-5, atol=2e-6, frac=0.99, what="tcgen05 preps + head: features vs oracle")
    assert_close_frac(all_xyz, want_xyz2, rtol=1e-5, atol=2e-6, frac=0.99, what="tcgen05 preps + head: xyz vs oracle")
