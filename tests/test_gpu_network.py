"""GPU parity of the per-point MLP / DenseEdgeConv / Level / Net mirrors against the CPU oracle
(oracle/ref_net.py, itself bit-identical to the unmodified reference Python on CPU).

Float features: 1e-5 relative (+1e-6 absolute for values near zero), the tolerance BASELINE.json states.
Where a stage contains a discrete choice (kNN membership at a near-tie, an FPS arg-max at a near-tie) the
downstream values can legitimately differ in a few places; those tests either inject the oracle's indices
(teacher forcing, strict tolerance everywhere) or state the share of elements that must agree."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_net
from tests.util import assert_close_frac, cloud_match_fraction

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def params():
    return ref_net.make_params(4, seed=1)


def _net(pu3, params, cuda, levels=4, knn=32):
    net = pu3.Net(max_up_ratio=2 ** levels, step_ratio=2, knn=knn, growth_rate=12, dense_n=3, fm_knn=5)
    sub = {k: v for k, v in params.items() if int(k.split(".")[1].split("_")[1]) <= levels}
    net.load_state_dict(sub, strict=True)
    return net.to(cuda)


@pytest.mark.parametrize("b,n,cin,cout,relu", [(3, 312, 3, 24, False), (2, 312, 84, 24, True), (5, 311, 204, 24, True),
                                               (2, 624, 264, 128, False), (3, 624, 128, 128, True),
                                               (2, 625, 128, 64, True), (4, 624, 64, 3, False), (1, 7, 5, 130, True)])
def test_pointwise_conv(pu3, cuda, b, n, cin, cout, relu):
    g = torch.Generator().manual_seed(cin * cout + n)
    x = torch.randn(b, cin, n, generator=g); w = torch.randn(cout, cin, 1, generator=g) * 0.2
    bias = torch.randn(cout, generator=g)
    want = F.conv1d(x.double(), w.double(), bias.double())
    want = F.relu(want) if relu else want
    got = pu3.fused.pointwise_conv(x.to(cuda), w.to(cuda), bias.to(cuda), relu=relu)
    assert_close_frac(got, want, rtol=1e-5, atol=1e-5, what="pointwise_conv")


def test_pointwise_conv_slices_and_residual(pu3, cuda):
    g = torch.Generator().manual_seed(5)
    buf = torch.randn(4, 100, 312, generator=g).to(cuda)
    out = torch.zeros(4, 50, 312, device=cuda)
    w = (torch.randn(24, 60, generator=g) * 0.1).to(cuda); bias = torch.randn(24, generator=g).to(cuda)
    pu3.fused.conv_into(buf[:, 40:], w, bias, out[:, 10:34], relu=True)
    want = F.relu(F.conv1d(buf[:, 40:].cpu().double(), w.cpu().double().unsqueeze(-1), bias.cpu().double()))
    assert_close_frac(out[:, 10:34], want, atol=1e-5)
    assert float(out[:, :10].abs().sum()) == 0 and float(out[:, 34:].abs().sum()) == 0
    res = torch.randn(4, 24, 156, generator=g).to(cuda)
    o2 = torch.empty(4, 24, 312, device=cuda)
    pu3.fused.conv_into(buf[:, 40:].contiguous(), w, bias, o2, residual=res, res_div=2)
    want2 = F.conv1d(buf[:, 40:].cpu().double(), w.cpu().double().unsqueeze(-1), bias.cpu().double()) + \
        res.cpu().double().repeat_interleave(2, dim=2)
    assert_close_frac(o2, want2, atol=1e-5)


@pytest.mark.parametrize("b,n,k", [(3, 312, 32), (2, 312, 16), (2, 100, 40), (1, 3000, 32)])
def test_dense_edge_conv_with_oracle_indices(pu3, cuda, params, b, n, k):
    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(b, 24, n, generator=g)
    pre = "levels.level_1.layer2"
    want, idx = ref_net.dense_edge_conv(params, pre, x, k, 3)
    ws = [params[f"{pre}.mlps.{i}.weight"].to(cuda) for i in range(3)]
    bs = [params[f"{pre}.mlps.{i}.bias"].to(cuda) for i in range(3)]
    with torch.no_grad():
        got, gidx = pu3.fused.dense_edge_conv(x.to(cuda), ws, bs, k, idx=idx.to(cuda))
    assert torch.equal(gidx.cpu(), idx)
    assert_close_frac(got, want, rtol=1e-5, atol=2e-6, what="dense_edge_conv")


def test_dense_edge_conv_own_knn(pu3, cuda, params):
    g = torch.Generator().manual_seed(2)
    x = torch.rand(4, 24, 312, generator=g)
    pre = "levels.level_2.layer3"
    want, idx = ref_net.dense_edge_conv(params, pre, x, 32, 3)
    ws = [params[f"{pre}.mlps.{i}.weight"].to(cuda) for i in range(3)]
    bs = [params[f"{pre}.mlps.{i}.bias"].to(cuda) for i in range(3)]
    with torch.no_grad():
        got, gidx = pu3.fused.dense_edge_conv(x.to(cuda), ws, bs, 32)
    assert gidx.shape == idx.shape and gidx.dtype == torch.int64
    assert (gidx.cpu() == idx).double().mean().item() > 0.9995
    # a flipped rank-32 neighbour at a near-tie changes one max() input: allow 0.1% of the outputs
    assert_close_frac(got, want, rtol=1e-5, atol=2e-6, frac=0.999, what="dense_edge_conv (own kNN)")


def test_dense_edge_conv_module_grad_path(pu3, cuda, params):
    # training: gradients flow (differentiable composition on our group_knn) and match the oracle's autograd
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(2, 24, 120, generator=g)
    pre = "levels.level_1.layer1"
    xr = x0.clone().requires_grad_()
    Pr = {k: v.clone().requires_grad_() for k, v in params.items() if k.startswith(pre)}
    yr, _ = ref_net.dense_edge_conv(Pr, pre, xr, 16, 3)
    yr.square().sum().backward()
    mod = pu3.layers.DenseEdgeConv(24, 12, 3, 16).to(cuda)
    mod.load_state_dict({k[len(pre) + 1:]: v for k, v in params.items() if k.startswith(pre)})
    xc = x0.clone().to(cuda).requires_grad_()
    y, _ = mod(xc)
    y.square().sum().backward()
    assert_close_frac(y, yr, rtol=1e-5, atol=2e-6, frac=0.999)
    assert_close_frac(xc.grad, xr.grad, rtol=1e-4, atol=1e-5, frac=0.995)
    assert_close_frac(mod.mlps[0].weight.grad, Pr[f"{pre}.mlps.0.weight"].grad, rtol=1e-3, atol=1e-4, frac=0.99)


def test_level_forward_first_level(pu3, cuda, params):
    net = _net(pu3, params, cuda).eval()
    g = torch.Generator().manual_seed(7)
    xyz = ref_net.normalize_point_batch(torch.rand(3, 3, 312, generator=g))[0]
    want_xyz, want_feat = ref_net.level_forward(params, "levels.level_1", xyz, xyz, None, knn=32)
    with torch.no_grad():
        got_xyz, got_feat = net.levels["level_1"](xyz.to(cuda), xyz.to(cuda))
    assert got_xyz.shape == (3, 3, 624) and got_feat.shape == (3, 264, 312)
    # near-tie kNN flips propagate through the later dense blocks: a small share of features may differ
    assert_close_frac(got_feat, want_feat, rtol=1e-5, atol=2e-6, frac=0.99, what="level features")
    assert_close_frac(got_xyz, want_xyz, rtol=1e-5, atol=2e-6, frac=0.99, what="level xyz")
    assert float((got_xyz.cpu() - want_xyz).abs().max()) < 1e-3


def test_level_forward_with_previous_level(pu3, cuda, params):
    net = _net(pu3, params, cuda).eval()
    g = torch.Generator().manual_seed(8)
    prev_xyz = torch.rand(2, 3, 624, generator=g)
    prev_feat = torch.randn(2, 264, 624, generator=g)
    # the current level's cloud: twice as dense, near (never on: h = 0 gives NaN weights in the reference,
    # upsampler.py:247-249) the previous level's points; 3 tiles per cloud, each the 312-NN of a seed
    cloud = prev_xyz.repeat(1, 1, 2) + 0.01 * torch.randn(2, 3, 1248, generator=g)
    seeds = cloud[:, :, :3].contiguous()
    tiles, _, _ = ref_net.group_knn(312, seeds, cloud, unique=False)
    tiles = tiles.permute(0, 2, 1, 3).reshape(6, 3, 312)
    tn = ref_net.normalize_point_batch(tiles)[0]
    wants = []
    for i in range(2):  # the reference handles one cloud per call and expand()s the previous level
        wants.append(ref_net.level_forward(params, "levels.level_2", tiles[3 * i:3 * i + 3], tn[3 * i:3 * i + 3],
                                           (prev_xyz[i:i + 1], prev_feat[i:i + 1]), knn=32))
    want_xyz = torch.cat([w[0] for w in wants]); want_feat = torch.cat([w[1] for w in wants])
    with torch.no_grad():
        got_xyz, got_feat = net.levels["level_2"](tiles.to(cuda), tn.to(cuda),
                                                  previous_level4=(prev_xyz.to(cuda), prev_feat.to(cuda)), group=3)
    assert_close_frac(got_feat, want_feat, rtol=1e-5, atol=5e-6, frac=0.99, what="level-2 features")
    assert_close_frac(got_xyz, want_xyz, rtol=1e-5, atol=5e-6, frac=0.99, what="level-2 xyz")


def test_net_eval_batched_equals_per_patch_calls(pu3, cuda, params):
    """BASELINE config 2 semantics at reduced depth: B patches in one call == B independent B=1 calls."""
    net = _net(pu3, params, cuda, levels=2).eval()
    g = torch.Generator().manual_seed(11)
    x = ref_net.normalize_point_batch(torch.rand(3, 3, 312, generator=g))[0]
    with torch.no_grad():
        together = net(x.to(cuda), ratio=4)
        alone = torch.cat([net(x[i:i + 1].to(cuda), ratio=4) for i in range(3)])
    assert together.shape == (3, 3, 1248)
    # not bit-equal yet: the torch reductions still used between kernels (normalisation, skip weights) pick
    # their summation order from the batch shape; 1e-7 noise can flip an FPS near-tie.  The clouds coincide.
    for i in range(3):
        assert cloud_match_fraction(together[i].cpu(), alone[i].cpu(), tol=1e-4) > 0.97


@pytest.mark.parametrize("ratio", [4, 16])
def test_net_eval_against_oracle(pu3, cuda, params, ratio):
    levels = {4: 2, 16: 4}[ratio]
    net = _net(pu3, params, cuda, levels=levels).eval()
    P = {k: v for k, v in params.items() if int(k.split(".")[1].split("_")[1]) <= levels}
    # End to end the result passes through FPS arg-max rounds and kNN selections whose near-ties may break
    # differently (1e-7 coordinate noise): one flipped seed moves a whole tile, so the share of points with an exact
    # twin is chaotic in the last few percent (profiles/e2e_share.py: 0.967 .. 1.000 over inputs and over
    # rounding-equivalent kernel variants).  Three inputs: the clouds must coincide almost everywhere on each, essentially
    # everywhere on most, and lie on the same surface sampling (Chamfer distance far below the point spacing) on all.
    shares = []
    for seed in (13, 14, 15):
        g = torch.Generator().manual_seed(seed)
        x = ref_net.normalize_point_batch(torch.rand(1, 3, 312, generator=g))[0]
        with torch.no_grad():
            want = ref_net.net_forward(P, x, ratio=ratio, max_up_ratio=ratio)
            got = net(x.to(cuda), ratio=ratio).cpu()
        assert got.shape == want.shape == (1, 3, 312 * ratio)
        shares.append(cloud_match_fraction(got[0], want[0], tol=1e-4))
        d = torch.cdist(got[0].t().double(), want[0].t().double())
        spacing = torch.cdist(want[0].t().double(), want[0].t().double()).topk(2, largest=False)[0][:, 1].mean()
        assert float(d.min(1)[0].mean()) < 0.02 * float(spacing) and float(d.min(0)[0].mean()) < 0.02 * float(spacing)
    assert min(shares) > 0.95 and sorted(shares)[1] > 0.99, shares


def test_net_train_forward_backward_against_oracle(pu3, cuda, params):
    levels, ratio, B = 2, 4, 2
    P = {k: v.clone().requires_grad_() for k, v in params.items() if int(k.split(".")[1].split("_")[1]) <= levels}
    g = torch.Generator().manual_seed(17)
    x = torch.rand(B, 3, 312, generator=g); gt = torch.rand(B, 3, 312 * ratio, generator=g)
    seeds = {2: torch.randint(0, 624, (B, 1), generator=g, dtype=torch.int32)}
    pr, gr = ref_net.net_forward(P, x, ratio=ratio, gt=gt, training=True, max_up_ratio=ratio, seed_idx_per_level=seeds)
    loss_r = ref_net.chamfer_loss(pr, gr)
    loss_r.backward()
    net = _net(pu3, params, cuda, levels=levels).train()
    pc, gc = net(x.to(cuda), ratio=ratio, gt=gt.to(cuda), seed_idx_per_level={2: seeds[2].to(cuda)})
    loss = pu3.ChamferLoss()(pc, gc)
    loss.backward()
    assert pc.shape == pr.shape == (B, 3, 624) and gc.shape == gr.shape
    # the gt patch is a pure gather of gt points around the seed; its ORDER (by distance to a seed that carries
    # 1e-7 noise) may differ at near-ties, the point set may not
    for i in range(B):
        assert cloud_match_fraction(gc[i].cpu(), gr[i], tol=0.0) > 0.995
    assert_close_frac(pc, pr, rtol=1e-5, atol=5e-6, frac=0.99, what="train prediction")
    assert abs(loss.item() - loss_r.item()) <= 1e-4 * abs(loss_r.item())
    gname = "levels.level_2.fc_layer2.conv.weight"
    got_g = dict(net.named_parameters())[gname].grad
    assert_close_frac(got_g, P[gname].grad, rtol=1e-3, atol=1e-5, frac=0.98, what="head weight grad")
    gname = "levels.level_1.layer0.conv.weight"
    got_g = dict(net.named_parameters())[gname].grad
    assert_close_frac(got_g, P[gname].grad, rtol=2e-2, atol=1e-4, frac=0.9, what="first-layer weight grad")


def test_fused_skip_connection_matches_composition(pu3, cuda, params):
    """csrc/skip.cu against the operator-by-operator skip connection (itself checked against the oracle above),
    on a ragged batch: 5 tiles belonging to 2 requests whose previous-level clouds have different sizes."""
    net = _net(pu3, params, cuda).eval()
    level = net.levels["level_3"]
    g = torch.Generator().manual_seed(23)
    sizes = [936, 624]
    prev_xyz = torch.zeros(2, 3, 936); prev_feat = torch.zeros(2, 264, 936)
    for i, n in enumerate(sizes):
        base = torch.rand(3, n // 3, generator=g)
        prev_xyz[i, :, :n] = base.repeat(1, 3) + (torch.arange(n) >= n // 3).float() * 0.0   # exact duplicates (overlapping tiles)
        prev_feat[i, :, :n] = torch.randn(264, n, generator=g)
    xyz = torch.rand(5, 3, 312, generator=g)
    x = torch.randn(5, 264, 312, generator=g)
    owner = torch.tensor([0, 0, 0, 1, 1], dtype=torch.int32, device=cuda)
    R = pu3.operations.Ragged(owner, owner, 2, n_arr=torch.tensor(sizes, dtype=torch.int32, device=cuda))
    with torch.no_grad():
        want = level._skip_connection(x.to(cuda), xyz.to(cuda), (prev_xyz.to(cuda), prev_feat.to(cuda)), None, R)
        got = level._skip_connection_fused(x.to(cuda).clone(), xyz.to(cuda),
                                           (prev_xyz.to(cuda), prev_feat.to(cuda).transpose(1, 2).contiguous()), None, R)
    assert_close_frac(got, want, rtol=1e-5, atol=2e-6, what="fused skip connection")
    # the compile-time (k=5, c=264) kernel and the runtime-shape kernel do the same arithmetic in the same order
    lib = pu3._lib.lib()
    try:
        lib.pu3_skip_force_generic(1)
        with torch.no_grad():
            gen = level._skip_connection_fused(x.to(cuda).clone(), xyz.to(cuda),
                                               (prev_xyz.to(cuda), prev_feat.to(cuda).transpose(1, 2).contiguous()), None, R)
    finally:
        lib.pu3_skip_force_generic(0)
    assert torch.equal(gen, got), "fixed-shape and generic skip kernels differ"
    # and the oracle itself, request by request (the reference expand()s one previous cloud per call)
    for lo, hi, c in ((0, 3, 0), (3, 5, 1)):
        n = sizes[c]
        nb_xyz, nb_idx, _ = ref_net.group_knn(5, xyz[lo:hi], prev_xyz[c:c + 1, :, :n].expand(hi - lo, -1, -1), unique=True)
        pf = prev_feat[c:c + 1, :, :n].expand(hi - lo, -1, -1).unsqueeze(2).expand(-1, -1, 312, -1)
        nb_feat = torch.gather(pf, 3, nb_idx.unsqueeze(1).expand(-1, 264, -1, -1))
        _, ws = ref_net.exponential_distance(xyz[lo:hi], nb_xyz)
        _, wf = ref_net.exponential_distance(x[lo:hi], nb_feat)
        w = ws * wf
        w = w / torch.sum(w + 1e-5, dim=-1, keepdim=True)
        ref = 0.2 * torch.sum(w * nb_feat, dim=-1) + x[lo:hi]
        assert_close_frac(got[lo:hi], ref, rtol=1e-5, atol=2e-6, frac=0.999, what="fused skip connection vs oracle")


def test_to_point_major(pu3, cuda):
    x = torch.randn(3, 70, 45, device=cuda)
    slot = torch.tensor([2, 0, 5], dtype=torch.int64, device=cuda)
    out = torch.zeros(6 * 45, 70, device=cuda)
    pu3._lib.launch("pu3_to_point_major_f32", x, 3, 70, 45, x.data_ptr(), slot.data_ptr(), out.data_ptr())
    for t, sl in enumerate(slot.tolist()):
        assert torch.equal(out[sl * 45:(sl + 1) * 45], x[t].t())


def test_pointwise_conv_fast_and_generic_kernels_agree(pu3, cuda):
    """Aligned inputs take the double-buffered FFMA2 kernel, everything else the generic one: same numbers
    (both accumulate over input channels in ascending order with fused multiply-adds)."""
    import ctypes
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    g = torch.Generator().manual_seed(77)
    for b, n, cin, cout, relu in [(3, 312, 84, 24, True), (2, 624, 264, 128, False), (2, 624, 128, 64, True),
                                  (2, 624, 64, 3, False), (5, 312, 3, 24, False), (1, 8, 17, 130, True)]:
        x = torch.randn(b, cin, n, generator=g).to(cuda); w = (torch.randn(cout, cin, generator=g) * 0.2).to(cuda)
        bias = torch.randn(cout, generator=g).to(cuda)
        outs = []
        for force in (0, 1):
            lib.pu3_pointwise_force_generic(force)
            try:
                o = torch.empty(b, cout, n, device=cuda)
                pu3.fused.conv_into(x, w, bias, o, relu=relu)
                outs.append(o)
            finally:
                lib.pu3_pointwise_force_generic(0)
        assert torch.equal(outs[0], outs[1]), (b, n, cin, cout)


def test_edgeconv_fast_and_generic_kernels_agree(pu3, cuda, params):
    """k <= 32 takes the FFMA2 two-edges-per-lane kernel.  Layers 1 and 2 accumulate in the generic kernel's order; layer 0
    is re-associated further (W0b n_j is computed once per point and gathered per edge instead of W0b (n_j - c) per edge),
    which changes its rounding by a few ulp of |W0b| |x| -- far inside the path's 1e-5 tolerance."""
    import ctypes
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    g = torch.Generator().manual_seed(5)
    pre = "levels.level_3.layer2"
    ws = [params[f"{pre}.mlps.{i}.weight"].to(cuda) for i in range(3)]
    bs = [params[f"{pre}.mlps.{i}.bias"].to(cuda) for i in range(3)]
    for b, n, k in [(3, 312, 32), (2, 100, 16), (1, 45, 31), (4, 312, 5)]:
        x = torch.randn(b, 24, n, generator=g).to(cuda)
        idx = torch.stack([torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(n)]) for _ in range(b)]).to(cuda)
        outs = []
        for force in (0, 1):
            lib.pu3_edgeconv_force_generic(force)
            try:
                with torch.no_grad():
                    outs.append(pu3.fused.dense_edge_conv(x, ws, bs, k, idx=idx)[0])
            finally:
                lib.pu3_edgeconv_force_generic(0)
        assert_close_frac(outs[0], outs[1], rtol=1e-5, atol=2e-6, what=f"fast vs generic edge-conv {(b, n, k)}")
        assert torch.equal(outs[0][:, 36:], outs[1][:, 36:])             # the pass-through channels are copies


def test_reference_import_names_resolve_through_the_shim(pu3, cuda):
    """What `main.py` of the reference imports (main.py:12-17, operations.py:2-6, model_loss.py:2) resolves to this
    package when 3pu_pytorch_b200/shim is first on sys.path."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch, sampling, losses, faiss\n"
            "from network import operations\n"
            "from network.upsampler import Net\n"
            "from network.model_loss import ChamferLoss\n"
            "from network.layers import Conv1d, Conv2d, DenseEdgeConv\n"
            "net = Net(max_up_ratio=4, step_ratio=2, knn=16, growth_rate=12, dense_n=3, fm_knn=5).cuda().eval()\n"
            "x = operations.normalize_point_batch(torch.rand(2, 3, 312).cuda())[0]\n"
            "with torch.no_grad(): y = net.forward(x, ratio=4)\n"
            "idx, pts = operations.furthest_point_sample(y, 100)\n"
            "print(tuple(y.shape), tuple(pts.shape), sampling.furthest_sampling.__module__)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "3pu_pytorch_b200", "shim"), root]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "(2, 3, 1248) (2, 3, 100) 3pu_pytorch_b200.sampling" in out.stdout


def test_whole_shape_pipeline_against_oracle(pu3, cuda, params):
    """BASELINE config 5 at reduced size: FPS seeds -> kNN patches -> batched upsample -> merge -> FPS."""
    levels, ratio, n_shape = 2, 4, 1000
    net = _net(pu3, params, cuda, levels=levels).eval()
    P = {k: v for k, v in params.items() if int(k.split(".")[1].split("_")[1]) <= levels}
    g = torch.Generator().manual_seed(31)
    pc = ref_net.normalize_point_batch(torch.rand(1, 3, n_shape, generator=g))[0]
    got = pu3.pipeline.upsample_shape(net, pc.to(cuda), num_point=312, patch_num_ratio=3, up_ratio=ratio).cpu()
    # the reference's loop (main.py:225-246, 375-380), on the oracle
    num_patches = int(n_shape / 312 * 3)
    _, seeds = ref_net.furthest_point_sample(pc, num_patches)
    patches, _, _ = ref_net.group_knn(312, seeds, pc)
    ups = []
    with torch.no_grad():
        for k in range(num_patches):
            patch, c, r = ref_net.normalize_point_batch(patches[:, :, k, :])
            ups.append(ref_net.net_forward(P, patch, ratio=ratio, max_up_ratio=ratio) * r + c)
    pred = torch.cat(ups, dim=-1)
    _, want = ref_net.furthest_point_sample(pred, n_shape * ratio)
    assert got.shape == want.shape == (1, 3, n_shape * ratio)
    assert cloud_match_fraction(got[0], want[0], tol=1e-4) > 0.95


def test_static_tile_slots_equal_the_synchronous_path(pu3, cuda, params):
    """Batched eval without host round trips (Net.static_tiles): requests whose outlier filter removes points get fewer
    tiles than the static slot count; the spare slots repeat the first tile and must not change anything."""
    net = _net(pu3, params, cuda).eval()
    level = net.levels["level_2"]
    g = torch.Generator().manual_seed(31)
    B, N, k = 4, 624, 312
    xyz = torch.rand(B, 3, N, generator=g)
    xyz[1, :, 5] += 40.0; xyz[1, :, 77] -= 35.0          # two outliers in request 1 -> 622 points -> 9 tiles instead of 10
    xyz[3, :, 600] += 50.0                               # one outlier in request 3
    old_xyz = torch.rand(B, 3, 312, generator=g)
    old_feat_pm = torch.randn(B, 312, 264, generator=g)
    old_n = torch.full((B,), 312, dtype=torch.int32, device=cuda)
    with torch.no_grad():
        want = net._eval_level_batched(level, xyz.to(cuda), old_xyz.to(cuda), old_feat_pm.to(cuda), old_n, k, 1248, True)
        bad = torch.zeros((), dtype=torch.int32, device=cuda)
        got = net._eval_level_static(level, xyz.to(cuda), old_xyz.to(cuda), old_feat_pm.to(cuda), old_n, k, 1248, True, bad)
    assert not bool(bad)
    assert want[3].tolist() == got[3].tolist()                                  # valid previous-level sizes: P_b * 312
    assert min(got[3].tolist()) < 3120 == max(got[3].tolist())                  # some requests use fewer tiles than slots
    assert torch.equal(want[0], got[0])                                         # the resampled clouds, bit for bit
    for b in range(B):
        nb = int(got[3][b])
        assert torch.equal(want[1][b, :, :nb], got[1][b, :, :nb]) and torch.equal(want[2][b, :nb], got[2][b, :nb])
    # a filtered cloud smaller than one tile is flagged (the caller then redoes the forward on the synchronous path).  With
    # the reference's filter (nearest-neighbour distance < 5 x mean) fewer than a fifth of a cloud can go, so this cannot
    # happen for clouds of >= 2 tiles: tests/test_gpu_glue.py exercises the flag on the kernel with a tile of 600 of 624 points.
    # whole forward: both paths give the same clouds
    x = ref_net.normalize_point_batch(torch.rand(3, 3, 312, generator=g))[0].to(cuda)
    with torch.no_grad():
        net.static_tiles = True
        a = net(x, ratio=16)                 # captured CUDA graph
        a2 = net(x, ratio=16)                # replay
        net.use_cuda_graph = False
        c = net(x, ratio=16)                 # the same launches, eager
        net.static_tiles = False
        b = net(x, ratio=16)
        net.static_tiles, net.use_cuda_graph = True, True
    assert torch.equal(a, b) and torch.equal(a, a2) and torch.equal(a, c)


def test_full_size_batch_properties(pu3, cuda, params):
    """BASELINE config 2 at full size (B=32, 312 -> 4992): properties that need no oracle run -- deterministic bit for
    bit across runs, finite, every request bit-identical to its own B=1 call (independent of its batch neighbours),
    upsampled points stay near the input's support."""
    net = _net(pu3, params, cuda).eval()
    g = torch.Generator().manual_seed(101)
    x = ref_net.normalize_point_batch(torch.rand(32, 3, 312, generator=g))[0].to(cuda)
    with torch.no_grad():
        a = net(x, ratio=16)
        b = net(x, ratio=16)
        solo = {i: net(x[i:i + 1], ratio=16) for i in (0, 17, 31)}
    assert a.shape == (32, 3, 4992) and bool(torch.isfinite(a).all())
    assert torch.equal(a, b)                                                   # no atomics / races on the eval path
    for i, s in solo.items():
        assert torch.equal(a[i], s[0]), i                                      # a request does not see its batch neighbours
    # with untrained (xavier) weights the residuals are large, but the clouds stay bounded and near their inputs
    assert float(a.norm(dim=1).max()) < 3.0
    d = torch.cdist(x.transpose(1, 2), a.transpose(1, 2)).min(dim=2)[0]          # (32,312)
    assert float(d.max()) < 1.0
