"""GPU: STRICT parity of the continuous stages.  End-to-end clouds can only be compared statistically: a neighbour or an FPS pick
at a near-tie legitimately flips on 1e-7 noise and the next level's re-tiling amplifies it (profiles/debug/dropin_divergence.py).
Here the discrete choices are taken from the oracle (neighbour lists injected through pu3_level_set_knn_override; every stage of
Net.forward started from the oracle's own intermediate state), and then EVERY element must agree to 1e-5 -- so a real error in
the layer / skip / head composition cannot hide inside a "99 % of the elements" allowance."""
import numpy as np
import pytest
import torch

from oracle import ref_net
from tests.util import cloud_match_fraction

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def params():
    return ref_net.make_params(4, seed=1)


def _net(pu3, params, cuda):
    net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
    net.load_state_dict(params, strict=True)
    return net.to(cuda).eval()


class _Override:
    """inject the oracle's neighbour lists into the level engine for the duration of a with-block"""

    def __init__(self, pu3, cuda, rec, slots=None):
        self.lib = pu3._lib.lib()

        def pad(t):     # static tile slots past the request's own tile count repeat its first tile (Net.static_tiles)
            if slots is not None and t.shape[0] < slots:
                t = torch.cat([t, t[:1].expand(slots - t.shape[0], -1, -1)], dim=0)
            return t
        self.knn = [pad(t).to(torch.int32).contiguous().to(cuda) for t in rec["knn"]]
        self.skip = pad(rec["skip"]).to(torch.int64).contiguous().to(cuda) if "skip" in rec else None

    def __enter__(self):
        self.lib.pu3_level_set_knn_override(*[t.data_ptr() for t in self.knn], self.skip.data_ptr() if self.skip is not None else None)

    def __exit__(self, *exc):
        self.lib.pu3_level_set_knn_override(None, None, None, None, None)
        return False


def _assert_all_close(got, want, what, rtol=1e-5, atol=2e-6):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    worst = float((err / bound).max())
    assert worst <= 1.0, f"{what}: an element is {worst:.2f}x over the 1e-5 bound (max abs err {float(err.max()):.3e}, scale {float(want.abs().max()):.3e})"


@pytest.mark.parametrize("with_prev", [False, True])
def test_level_forward_teacher_forced_every_element(pu3, cuda, params, with_prev):
    net = _net(pu3, params, cuda)
    g = torch.Generator().manual_seed(5 + with_prev)
    T, N = 6, 312
    xyz = torch.rand(T, 3, N, generator=g) * 0.3 + 0.2
    xn = ref_net.normalize_point_batch(xyz)[0]
    prev = None
    if with_prev:
        prev_xyz = torch.rand(T, 3, 624, generator=g) * 0.3 + 0.2
        prev_feat = torch.randn(T, 264, 624, generator=g)
        prev = (prev_xyz, prev_feat)
    name = "level_3" if with_prev else "level_1"
    rec = {}
    with torch.no_grad():
        want_xyz, want_feat = ref_net.level_forward(params, f"levels.{name}", xyz, xn, prev, knn=32, record=rec)
    level = net.levels[name]
    kw = {}
    if with_prev:
        kw = dict(previous_level4=(prev_xyz.to(cuda), prev_feat.transpose(1, 2).contiguous().to(cuda)), prev_point_major=True)
    with torch.no_grad(), _Override(pu3, cuda, rec):
        got_xyz, got_feat = level(xyz.to(cuda), xn.to(cuda), **kw)
        torch.cuda.synchronize()
    _assert_all_close(got_feat, want_feat, f"{name} features")
    _assert_all_close(got_xyz, want_xyz, f"{name} coordinates")


def test_net_eval_stage_by_stage_from_the_oracle_state(pu3, cuda, params):
    """Every level past the first, started from the ORACLE's state before that level (upsampler.py:128-159): outlier filter, FPS
    seeds and kNN tiles must reproduce the oracle's tiles exactly; with its neighbour lists injected, the merged cloud and the
    features handed to the next level agree on every element; the resampled cloud agrees as a set."""
    net = _net(pu3, params, cuda)
    g = torch.Generator().manual_seed(3)
    x = ref_net.normalize_point_batch(torch.rand(1, 3, 312, generator=g))[0]
    trace = {}
    with torch.no_grad():
        ref_net.net_forward(params, x, ratio=16, max_up_ratio=16, knn=32, trace=trace)
    for l in (2, 3, 4):
        st = trace[l]
        n_in, No = st["xyz_in"].shape[2], st["old_xyz"].shape[2]
        old_n = torch.full((1,), No, dtype=torch.int32, device=cuda)
        bad = torch.zeros((), dtype=torch.int32, device=cuda)
        dbg = {}
        slots = int(n_in / 312 * 5)
        with torch.no_grad(), _Override(pu3, cuda, st, slots=slots):
            out, prev_xyz, feat_pm, pk = net._eval_level_static(
                net.levels[f"level_{l}"], st["xyz_in"].to(cuda), st["old_xyz"].to(cuda),
                st["old_feat"].transpose(1, 2).contiguous().to(cuda), old_n, 312, 312 * 2 ** l, True, bad, debug=dbg)
            torch.cuda.synchronize()
        assert int(bad) == 0
        P = st["patch"].shape[0]
        assert int(dbg["p_arr"][0]) == P <= slots == dbg["patch_xyz"].shape[0]        # same number of tiles (:76)
        assert torch.equal(dbg["patch_xyz"][:P].cpu(), st["patch"]), f"level {l}: tiles differ from the oracle's"
        merged = dbg["merged_pm"].transpose(1, 2)[:, :, :P * 624]                     # (1,3,P*624) valid part
        _assert_all_close(merged, st["merged"], f"level {l} merged cloud")
        _assert_all_close(feat_pm.transpose(1, 2)[:, :, :P * 312], st["feat"], f"level {l} features for the next level")
        assert torch.equal(prev_xyz[:, :, :P * 312].cpu(), st["next_old_xyz"])
        assert int(pk[0]) == P * 312
        assert cloud_match_fraction(out[0].cpu(), st["xyz_out"][0], tol=1e-5) > 0.99, f"level {l} resampled cloud"


@pytest.mark.parametrize("with_prev", [False, True])
def test_level_backward_teacher_forced_against_oracle_autograd(pu3, cuda, params, with_prev):
    """Train-mode Level as one native autograd node (level_train.py) against torch autograd of the ORACLE's Level (the reference's
    graph, upsampler.py:272-374), with the oracle's neighbour lists injected so that both differentiate the same graph: the
    gradient of a random linear functional of (coordinates, features) with respect to all 40 parameters, the normalised input
    cloud and the previous level's features.  Remaining discrete difference: the ReLU mask of an activation within rounding of
    zero (more of them after the skip connection, whose exp() weights move the features by ~1e-6).  ONE flipped unit at one of the
    ~2500 points moves a whole row of that layer's weight gradient and one entry of its bias gradient by that point's share,
    ~4e-4 .. 1e-3 of the tensor's scale -- that is the quantum below which a per-entry bound says nothing.  Bar for every
    gradient: relative L2 error <= 1e-3, 99.9 % of the entries within 1e-3 of scale (a bias-sized tensor: at most one entry
    beyond), none off by more than 1 % of scale."""
    name = "level_3" if with_prev else "level_1"
    g = torch.Generator().manual_seed(21 + with_prev)
    T, N = 4, 312
    xyz = torch.rand(T, 3, N, generator=g) * 0.3 + 0.2
    xn = ref_net.normalize_point_batch(xyz)[0]
    prev = None
    if with_prev:
        prev = (torch.rand(T, 3, 312, generator=g) * 0.3 + 0.2, torch.randn(T, 264, 312, generator=g))
    w_xyz = torch.randn(T, 3, 2 * N, generator=g)
    w_feat = torch.randn(T, 264, N, generator=g) * 0.1
    # ---- oracle: torch autograd on the reference graph
    P = {k: v.clone().requires_grad_() for k, v in params.items() if k.startswith(f"levels.{name}.")}
    xn_r = xn.clone().requires_grad_()
    prev_r = None if prev is None else (prev[0], prev[1].clone().requires_grad_())
    rec = {}
    oxyz, ofeat = ref_net.level_forward(P, f"levels.{name}", xyz, xn_r, prev_r, knn=32, training=True, record=rec)
    ((oxyz * w_xyz).sum() + (ofeat * w_feat).sum()).backward()
    # ---- ours: the native node, same neighbour lists
    net = _net(pu3, params, cuda).train()
    level = net.levels[name]
    xn_g = xn.to(cuda).requires_grad_()
    prev_g = None if prev is None else (prev[0].to(cuda), prev[1].to(cuda).requires_grad_())
    with _Override(pu3, cuda, rec):
        gxyz, gfeat = level(xyz.to(cuda), xn_g, previous_level4=prev_g)
        ((gxyz * w_xyz.to(cuda)).sum() + (gfeat * w_feat.to(cuda)).sum()).backward()
        torch.cuda.synchronize()
    _assert_all_close(gxyz, oxyz, f"{name} train forward coordinates")
    _assert_all_close(gfeat, ofeat, f"{name} train forward features")

    def close(got, want, what):
        got, want = got.detach().cpu().double(), want.detach().double()
        scale = float(want.abs().max()) + 1e-30
        err = (got - want).abs()
        frac = float((err <= 1e-3 * scale).double().mean())
        n_off = int((err > 1e-3 * scale).sum())
        l2 = float(err.norm() / (want.norm() + 1e-30))
        assert (frac >= 0.999 or n_off <= 1) and float(err.max()) <= 1e-2 * scale and l2 <= 1e-3, \
            f"{what}: {frac:.5f} of the entries within 1e-3 of scale {scale:.3e}, max err {float(err.max()):.3e}, rel L2 {l2:.2e}"
    close(xn_g.grad, xn_r.grad, "d / d xyz_normalized")
    if with_prev:
        close(prev_g[1].grad, prev_r[1].grad, "d / d previous features")
    got = dict(net.named_parameters())
    assert len(P) == 40
    for k in sorted(P):
        close(got[k].grad, P[k].grad, k)
