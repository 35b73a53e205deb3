"""CPU restatement (torch fp32, CPU tensors) of the reference's PYTHON hot path.

TEST INFRASTRUCTURE ONLY.  The product package never imports this file.  It is
the checker on the GPU box, where /root/reference does not exist, and the
"port" CPU baseline that bench.py times.

What is restated (reference file:line in each docstring):
  network/operations.py   group_knn, normalize_point_batch, furthest_point_sample, gather_points
  network/layers.py       DenseEdgeConv (Conv1d/Conv2d are plain 1x1 convolutions)
  network/upsampler.py    Level.forward, Net.forward (train zoom + eval tiling), extract_xyz_feature_patch
  network/model_loss.py   NmDistanceFunction (with the backward NameError at :22-23 removed), ChamferLoss
The CUDA-only pieces (FPS, gather, nmdistance) go through oracle_c.c.

Style: functional -- parameters come in as a flat dict with the reference's
state_dict key names, so one set of weights drives the
reference, this oracle and the product.

Parity pin: tests/golden/ref_py.npz was minted from the UNMODIFIED reference Python
(imported on CPU by oracle/reference_loader.py) by tests/golden/make_golden_py.py;
tests/test_cpu_ref_net_golden.py replays the stored inputs through this file and
requires bit-identical results (group_knn with duplicates and both layouts, one Level,
a 2-level eval forward, a train-mode forward with Chamfer loss and gradients, the
threshold branch of ChamferLoss), and re-derives the fixture from the reference where
/root/reference exists.
"""
from math import log

import numpy as np
import torch
import torch.nn.functional as F

from . import c_oracle


# --------------------------------------------------------------------------------------
# operations.py
# --------------------------------------------------------------------------------------
def normalize_point_batch(pc, NCHW=True):
    """operations.py:12-30: subtract the centroid, divide by the largest norm."""
    pdim, cdim = (2, 1) if NCHW else (1, 2)
    centroid = pc.mean(dim=pdim, keepdim=True)
    pc = pc - centroid
    radius = torch.sqrt((pc ** 2).sum(dim=cdim, keepdim=True)).max(dim=pdim, keepdim=True)[0]
    return pc / radius, centroid, radius


def pairwise_sqdist_expanded(q, p):
    """operations.py:151-162: |q|^2 - 2 q.p^T + |p|^2 with q (B,M,C), p (B,N,C) -> (B,M,N)."""
    rq = (q * q).sum(dim=2, keepdim=True)
    rp = (p * p).sum(dim=2, keepdim=True)
    inner = torch.matmul(q, p.permute(0, 2, 1))
    return rq - 2 * inner + rp.permute(0, 2, 1)


def duplicate_mask(points_bnc):
    """operations.py:192-200: 1 for every row that is NOT the first occurrence of its value
    (np.unique(..., axis=0, return_index=True) keeps first occurrences), shape (B,1,N) float."""
    pts = points_bnc.detach().cpu().numpy()
    B, N, _ = pts.shape
    dup = np.ones((B, 1, N), dtype=np.int32)
    for b in range(B):
        _, first = np.unique(pts[b], return_index=True, axis=0)
        dup[b, :, first] = 0
    return torch.from_numpy(dup).to(dtype=torch.float32)


def group_knn(k, query, points, unique=True, NCHW=True):
    """operations.py:165-216.  Returns (neighbours (B,C,M,k) [NCHW] or (B,M,k,C), idx (B,M,k) int64,
    dist (B,M,k) ascending).  Non-first duplicate points are pushed back by max(D) over the whole
    batch when unique (operations.py:204)."""
    if NCHW:
        p = points.transpose(2, 1).contiguous()
        q = query.transpose(2, 1).contiguous()
    else:
        p = points.contiguous()
        q = query.contiguous()
    assert p.size(1) >= k, "points size must be greater or equal to k"
    D = pairwise_sqdist_expanded(q, p)
    if unique:
        D = D + torch.max(D) * duplicate_mask(p).to(D.device)
    neg, idx = torch.topk(-D, k, dim=-1, sorted=True)
    C = p.size(-1)
    nb = torch.gather(p.unsqueeze(1).expand(-1, q.size(1), -1, -1), 2,
                      idx.unsqueeze(-1).expand(-1, -1, -1, C))
    if NCHW:
        nb = nb.permute(0, 3, 1, 2)
    return nb, idx, -neg


class _Gather(torch.autograd.Function):
    """operations.py:219-263 GatherFunction over sampling_cuda.cu:26-80."""

    @staticmethod
    def forward(ctx, features, idx):
        f = features.contiguous()
        i32 = idx.contiguous().to(torch.int32)
        out = torch.from_numpy(c_oracle.gather_fwd(f.detach().numpy(), i32.numpy()))
        ctx.save_for_backward(i32)
        ctx.n = f.size(2)
        return out

    @staticmethod
    def backward(ctx, g):
        (i32,) = ctx.saved_tensors
        gp = c_oracle.gather_bwd(g.contiguous().numpy(), i32.numpy(), ctx.n)
        return torch.from_numpy(gp), None


def gather_points(features, idx):
    return _Gather.apply(features, idx)


def furthest_point_sample(xyz, npoint, NCHW=True):
    """operations.py:303-323: FPS indices (B,npoint) int32 and the gathered coordinates."""
    pts = xyz.transpose(2, 1).contiguous() if NCHW else xyz.contiguous()
    idx = torch.from_numpy(c_oracle.fps(pts.detach().numpy(), int(npoint)))
    sampled = gather_points(pts.transpose(2, 1).contiguous(), idx)
    if not NCHW:
        sampled = sampled.transpose(2, 1).contiguous()
    return idx, sampled


# --------------------------------------------------------------------------------------
# model_loss.py
# --------------------------------------------------------------------------------------
class _NmDistance(torch.autograd.Function):
    """model_loss.py:5-28 over nmdistance_cuda.cu (the undefined d_dist1/d_dist2 lines dropped)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, i1, d2, i2 = c_oracle.nmdist_fwd(xyz1.detach().numpy(), xyz2.detach().numpy())
        d1, i1, d2, i2 = map(torch.from_numpy, (d1, i1, d2, i2))
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, i1, d2, i2

    @staticmethod
    def backward(ctx, g1, _a, g2, _b):
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        gx1, gx2 = c_oracle.nmdist_bwd(xyz1.detach().numpy(), xyz2.detach().numpy(),
                                       g1.contiguous().numpy(), g2.contiguous().numpy(), i1.numpy(), i2.numpy())
        return torch.from_numpy(gx1), torch.from_numpy(gx2)


def nndistance(xyz1, xyz2):
    return _NmDistance.apply(xyz1, xyz2)


def chamfer_loss(pred, gt, threshold=None, forward_weight=1.0):
    """model_loss.py:50-85 ChamferLoss.forward; accepts (B,3,N) or (B,N,3)."""
    assert pred.dim() == 3 and gt.dim() == 3
    if pred.size(2) != 3:
        pred = pred.transpose(2, 1).contiguous()
    if gt.size(2) != 3:
        gt = gt.transpose(2, 1).contiguous()
    p2g, _, g2p, _ = nndistance(pred, gt)
    if threshold is not None:
        ft = p2g.mean(dim=1, keepdim=True) * threshold
        bt = g2p.mean(dim=1, keepdim=True) * threshold
        p2g = torch.where(p2g < ft, p2g, torch.zeros_like(p2g))
        g2p = torch.where(g2p < bt, g2p, torch.zeros_like(g2p))
    cd = forward_weight * p2g.mean(dim=1) + g2p.mean(dim=1)
    return cd.mean()


# --------------------------------------------------------------------------------------
# layers.py / upsampler.py
# --------------------------------------------------------------------------------------
def _conv(x, w, b):
    """1x1 convolution exactly as nn.Conv1d / nn.Conv2d would run it."""
    return F.conv2d(x, w, b) if w.dim() == 4 else F.conv1d(x, w, b)


def dense_edge_conv(P, prefix, x, k, n=3, record=None):
    """layers.py:44-64 DenseEdgeConv.forward with idx=None.  x (B,C,N) -> y (B,C+n*growth,N), idx (B,N,k).
    record (a list, tests only): receives the full (B,N,k+1) neighbour lists incl. the dropped rank 0 (teacher forcing)."""
    nb, idx, _ = group_knn(k + 1, x, x, unique=True)          # layers.py:33
    if record is not None:
        record.append(idx.clone())
    idx = idx[:, :, 1:]                                       # drop rank 0, not "self" (layers.py:34-35)
    nb = nb[:, :, :, 1:]
    centre = x.unsqueeze(-1).expand_as(nb)
    y = torch.cat([centre, nb - centre], dim=1)               # layers.py:40-41
    for i in range(n):
        w, b = P[f"{prefix}.mlps.{i}.weight"], P[f"{prefix}.mlps.{i}.bias"]
        if i == 0:
            xr = x.unsqueeze(-1).repeat(1, 1, 1, k)
            y = torch.cat([F.relu(_conv(y, w, b)), xr], dim=1)
        elif i == n - 1:
            y = torch.cat([_conv(y, w, b), y], dim=1)         # last mlp: no ReLU (layers.py:58-59)
        else:
            y = torch.cat([F.relu(_conv(y, w, b)), y], dim=1)
    return y.max(dim=-1)[0], idx


def exponential_distance(points, nbrs):
    """upsampler.py:232-250: squared distance to each neighbour and exp(-d/(h/2)), h = mean_N(min_K d)."""
    if points.dim() == 3:
        points = points.unsqueeze(-1)
    d = ((points - nbrs) ** 2).sum(dim=1, keepdim=True).detach()
    h = d.min(dim=-1, keepdim=True)[0].mean(dim=-2, keepdim=True)
    return d, torch.exp(-d / (h / 2)).detach()


def level_forward(P, prefix, xyz, xyz_normalized, previous_level4=None, knn=32, fm_knn=5,
                  step_ratio=2, dense_n=3, training=False, record=None):
    """upsampler.py:272-374 Level.forward.  Returns (xyz' (B,3,N*r) normalised frame, features (B,264,N)).
    record (a dict, tests only): receives "knn" = the four blocks' neighbour lists and "skip" = the skip connection's."""
    knn_rec = None
    if record is not None:
        knn_rec = record.setdefault("knn", [])
    B, _, N = xyz_normalized.shape
    g = lambda name: (P[f"{prefix}.{name}.weight"], P[f"{prefix}.{name}.bias"])
    x = _conv(xyz_normalized.unsqueeze(-1), *g("layer0.conv")).squeeze(-1)
    y, _ = dense_edge_conv(P, f"{prefix}.layer1", x, knn, dense_n, knn_rec)
    x = torch.cat([y, x], dim=1)
    for li in (2, 3, 4):
        h = F.relu(_conv(x, *g(f"layer{li}_prep.conv")))
        y, _ = dense_edge_conv(P, f"{prefix}.layer{li}", h, knn, dense_n, knn_rec)
        x = torch.cat([y, x], dim=1)

    if previous_level4 is not None and fm_knn > 0:               # upsampler.py:317-347
        pxyz, pfeat = previous_level4
        if not training and pxyz.shape[0] != x.shape[0]:
            pxyz = pxyz.expand(B, -1, -1)
            pfeat = pfeat.expand(B, -1, -1)
        nb_xyz, nb_idx, _ = group_knn(fm_knn, xyz, pxyz, unique=True, NCHW=True)
        if record is not None:
            record["skip"] = nb_idx.clone()
        pf = pfeat.unsqueeze(2).expand(-1, -1, N, -1)
        nb_feat = torch.gather(pf, 3, nb_idx.unsqueeze(1).expand(-1, pf.size(1), -1, -1))
        _, ws = exponential_distance(xyz, nb_xyz)
        _, wf = exponential_distance(x, nb_feat)
        w = ws * wf
        w = w / torch.sum(w + 1e-5, dim=-1, keepdim=True)
        x = 0.2 * torch.sum(w * nb_feat, dim=-1) + x

    feats = x
    r = step_ratio
    assert r < 4, "gen_grid (step_ratio>=4) is not on the benchmarked path"
    code = torch.linspace(-0.2, 0.2, r).view(1, 1, r).repeat(B, 1, N)      # upsampler.py:264-270,352
    x = x.unsqueeze(-1).expand(-1, -1, -1, r).reshape(B, x.size(1), N * r).contiguous()
    x = torch.cat([x, code], dim=1).unsqueeze(-1)
    x = F.relu(_conv(x, *g("up_layer.up_layer1.conv")))
    x = F.relu(_conv(x, *g("up_layer.up_layer2.conv")))
    x = F.relu(_conv(x, *g("fc_layer1.conv")))
    x = _conv(x, *g("fc_layer2.conv")).squeeze(-1)
    x = x + xyz_normalized.unsqueeze(3).repeat(1, 1, 1, r).reshape(B, 3, N * r)
    return x, feats


def extract_patches(xyz, k, training, gt_xyz=None, gt_k=None, seed_idx=None):
    """upsampler.py:39-105 extract_xyz_feature_patch.
    train: one random seed per sample (seed_idx (B,1) int32 may be injected for reproducibility);
    eval : batch 1, outlier filter, FPS seeds, int(N'/k*5) patches."""
    B, _, N = xyz.shape
    if training:
        if seed_idx is None:
            seed_idx = torch.randint(0, N, (B, 1), dtype=torch.int32)
        seeds = gather_points(xyz, seed_idx)
    else:
        assert B == 1
        _, _, d = group_knn(2, xyz, xyz, unique=False, NCHW=True)
        d = d[:, :, 1]
        mask = d < 5 * d.mean(dim=1, keepdim=True)
        xyz = torch.masked_select(xyz, mask.unsqueeze(1).expand_as(xyz)).view(1, 3, -1)
        N = xyz.size(2)
        _, seeds = furthest_point_sample(xyz, int(N / k * 5))
        k = min(k, N)
    patches, _, _ = group_knn(k, seeds, xyz, unique=False, NCHW=True)
    patches = torch.cat(torch.unbind(patches, dim=2), dim=0)
    if gt_xyz is not None and gt_k is not None:
        gt, _, _ = group_knn(gt_k, seeds, gt_xyz, unique=False)
        gt = torch.cat(torch.unbind(gt, dim=2), dim=0)
    else:
        gt = None
    return patches, gt


def net_forward(P, xyz, ratio=16, gt=None, training=False, max_up_ratio=16, step_ratio=2, knn=32,
                fm_knn=5, dense_n=3, max_num_point=312, seed_idx_per_level=None, trace=None):
    """upsampler.py:107-189 Net.forward.  Note Net builds Level without fm_knn (upsampler.py:25-26),
    so the effective fm_knn is Level's default 5 whatever Net was given.
    trace (a dict, tests only): trace[l] receives the state around level l -- the cloud entering it, the previous level's
    (xyz, features) it reads, its tiles, the neighbour lists its Level used, the merged cloud before and after the
    resampling FPS -- for stage-wise teacher forcing."""
    B, _, N = xyz.shape
    levels = int(log(ratio, step_ratio))
    cap = min(N, max_num_point)
    old_xyz = old_feat = None
    for l in range(1, levels + 1):
        cur = step_ratio ** l
        pre = f"levels.level_{l}"
        if l == 1:
            old_xyz = xyz
            xyz, feat = level_forward(P, pre, xyz, xyz, None, knn, fm_knn, step_ratio, dense_n, training)
            old_feat = feat
            continue
        if xyz.size(-1) > cap:
            gt_k = cap * ratio // cur * step_ratio
            sidx = None if seed_idx_per_level is None else seed_idx_per_level.get(l)
            patch, gt = extract_patches(xyz, cap, training, gt_xyz=gt, gt_k=gt_k, seed_idx=sidx)
        else:
            patch = xyz
        patch_n, centroid, radius = normalize_point_batch(patch, NCHW=True)
        rec = None
        if trace is not None:
            rec = trace[l] = {"xyz_in": xyz.clone(), "old_xyz": old_xyz.clone(), "old_feat": old_feat.clone(),
                              "patch": patch.clone()}
        xyz, feat = level_forward(P, pre, patch, patch_n, (old_xyz, old_feat), knn, fm_knn, step_ratio,
                                  dense_n, training, record=rec)
        xyz = xyz * radius + centroid
        old_xyz, old_feat = patch, feat
        if not training and patch.shape[0] != B:                  # merge tiles, resample (upsampler.py:149-159)
            xyz = torch.cat(torch.split(xyz, B, dim=0), dim=2)
            old_xyz = torch.cat(torch.split(old_xyz, B, dim=0), dim=2)
            old_feat = torch.cat(torch.split(old_feat, B, dim=0), dim=2)
            if rec is not None:
                rec["merged"] = xyz.clone()
            _, xyz = furthest_point_sample(xyz, N * cur)
        if rec is not None:
            rec["xyz_out"] = xyz.clone(); rec["feat"] = old_feat.clone(); rec["next_old_xyz"] = old_xyz.clone()
    return (xyz, gt) if training else xyz


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def level_param_shapes():
    """(suffix, weight shape) of one Level; biases are (shape[0],)."""
    out = [("layer0.conv", (24, 3, 1, 1))]
    for li in (1, 2, 3, 4):
        out += [(f"layer{li}.mlps.0", (12, 48, 1, 1)), (f"layer{li}.mlps.1", (12, 36, 1, 1)),
                (f"layer{li}.mlps.2", (12, 48, 1, 1))]
    out += [("layer2_prep.conv", (24, 84, 1)), ("layer3_prep.conv", (24, 144, 1)), ("layer4_prep.conv", (24, 204, 1)),
            ("up_layer.up_layer1.conv", (128, 265, 1, 1)), ("up_layer.up_layer2.conv", (128, 128, 1, 1)),
            ("fc_layer1.conv", (64, 128, 1, 1)), ("fc_layer2.conv", (3, 64, 1, 1))]
    return out


def make_params(num_levels=4, seed=0, bias_scale=0.05):
    """Deterministic synthetic weights with the reference's state_dict keys.  Xavier-uniform bounds
    like upsampler.py:27-31, but biases are small non-zero values so that bias handling is exercised
    (a trained checkpoint has non-zero biases).  numpy PCG64 -> identical on every host."""
    rng = np.random.Generator(np.random.PCG64(seed))
    P = {}
    for l in range(1, num_levels + 1):
        for suffix, shape in level_param_shapes():
            fan_out, fan_in = shape[0], shape[1]
            bound = float(np.sqrt(6.0 / (fan_in + fan_out)))
            P[f"levels.level_{l}.{suffix}.weight"] = torch.from_numpy(
                rng.uniform(-bound, bound, size=shape).astype(np.float32))
            P[f"levels.level_{l}.{suffix}.bias"] = torch.from_numpy(
                rng.uniform(-bias_scale, bias_scale, size=(shape[0],)).astype(np.float32))
    return P
