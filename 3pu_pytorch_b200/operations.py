"""Operator layer: same names, signatures and return conventions as the reference's
network/operations.py (group_knn :165, gather_points :266, furthest_point_sample :303,
normalize_point_batch :12), running on libpu3_b200's sm_100a kernels.

CUDA tensors only.  A CPU tensor raises: this package has no CPU path by design (the CPU
restatement lives in oracle/ and is test infrastructure).
"""
import torch

from . import _lib
from . import sampling


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required (3pu_pytorch_b200 has no CPU fallback)")


# ------------------------------------------------------------------------------------------------
# normalize_point_batch
# ------------------------------------------------------------------------------------------------
def normalize_point_batch(pc, NCHW=True):
    """operations.py:12-30.  Returns (normalised pc, centroid, furthest_distance).
    CUDA float32 without autograd: one kernel (pu3_normalize_f32).  When a gradient is wanted (the train-mode zoom,
    upsampler.py:138) the differentiable composition of the same arithmetic runs instead."""
    if pc.is_cuda and pc.dtype == torch.float32 and pc.dim() == 3 and not (torch.is_grad_enabled() and pc.requires_grad):
        src = pc.contiguous()
        B = src.shape[0]
        n = src.shape[2] if NCHW else src.shape[1]
        if (src.shape[1] if NCHW else src.shape[2]) != 3:
            raise RuntimeError("normalize_point_batch: 3-D points expected")
        out = torch.empty_like(src)
        centroid = torch.empty((B, 3, 1) if NCHW else (B, 1, 3), dtype=torch.float32, device=src.device)
        radius = torch.empty(B, 1, 1, dtype=torch.float32, device=src.device)
        _lib.launch("pu3_normalize_f32", src, B, n, int(bool(NCHW)), src.data_ptr(), out.data_ptr(), centroid.data_ptr(),
                    radius.data_ptr())
        return out, centroid, radius
    point_axis, dim_axis = (2, 1) if NCHW else (1, 2)
    centroid = torch.mean(pc, dim=point_axis, keepdim=True)
    pc = pc - centroid
    furthest_distance, _ = torch.max(
        torch.sqrt(torch.sum(pc ** 2, dim=dim_axis, keepdim=True)), dim=point_axis, keepdim=True)
    pc = pc / furthest_distance
    return pc, centroid, furthest_distance


# ------------------------------------------------------------------------------------------------
# group_knn
# ------------------------------------------------------------------------------------------------
class Ragged:
    """Description of a ragged batch for group_knn (batched eval): batch element i reads cloud owner[i], which
    holds n_arr[owner[i]] valid points; it has m_arr[i] valid queries and belongs to duplicate-penalty group
    group_of[i].  All tensors int32 on the device; n_arr / m_arr may be None (= full rows)."""

    def __init__(self, owner, group_of, groups, n_arr=None, m_arr=None):
        self.owner, self.group_of, self.groups, self.n_arr, self.m_arr = owner, group_of, int(groups), n_arr, m_arr


def _dup_prepass_kernels(unique, C, N, k):
    """launches of the duplicate pre-pass (hash table, group flags, max D): none when the tiled feature-space kernel runs (it
    finds duplicates itself, csrc/group_knn.cu) -- n <= 320 and not one of the other specialised kernels"""
    if not unique:
        return 0
    feat_path = N <= 320 and k <= 64 and not (C == 3 and k <= 8)
    return 0 if feat_path else 3


def _knn_raw(k, q, p, unique, max_group, want_knn=True, want_dist=True, idx_dtype=torch.int64, ragged=None, set_order=False):
    """q (B,C,M), p (B/p_div,C,N) contiguous f32 on the same device -> (knn|None, idx, dist|None).
    With `ragged`, p is (clouds,C,N) and the mapping comes from the Ragged description; outputs of invalid
    rows / columns are zero.  set_order (indices-only int32 calls): rank 0 exact, ranks 1..k-1 as an unordered set
    (PU3_KNN_SET_ORDER of include/pu3_b200.h) -- what the fused DenseEdgeConv needs."""
    uflag = int(bool(unique)) | (2 if set_order else 0)
    B, C, M = q.shape
    Bp, _, N = p.shape
    if ragged is not None:
        dev = q.device
        alloc = torch.zeros
        knn = alloc(B, C, M, k, dtype=torch.float32, device=dev) if want_knn else None
        idx = alloc(B, M, k, dtype=idx_dtype, device=dev)
        dist = alloc(B, M, k, dtype=torch.float32, device=dev) if want_dist else None
        ws_bytes = _lib.lib().pu3_group_knn_workspace(B, C, M, N, k, 1, int(bool(unique)))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        _lib.launch("pu3_group_knn_ragged_f32", q, B, C, M, N, k, Bp, ragged.groups, _lib.ptr(ragged.owner),
                    _lib.ptr(ragged.group_of), _lib.ptr(ragged.n_arr), _lib.ptr(ragged.m_arr), _lib.ptr(q), _lib.ptr(p),
                    uflag, _lib.ptr(knn), _lib.ptr(idx if idx_dtype == torch.int64 else None),
                    _lib.ptr(idx if idx_dtype == torch.int32 else None), _lib.ptr(dist), _lib.ptr(ws), ws_bytes,
                    extra_kernels=_dup_prepass_kernels(unique, C, N, k),
                    tag=f"pu3_group_knn_f32[c={C},k={k},n<={N}]")
        return knn, idx, dist
    if Bp == 0 or B % Bp != 0:
        raise RuntimeError(f"group_knn: points batch {Bp} must divide query batch {B}")
    p_div = B // Bp
    if N < k:
        raise AssertionError("points size must be greater or equal to k")  # operations.py:186
    dev = q.device
    knn = torch.empty(B, C, M, k, dtype=torch.float32, device=dev) if want_knn else None
    idx = torch.empty(B, M, k, dtype=idx_dtype, device=dev)
    dist = torch.empty(B, M, k, dtype=torch.float32, device=dev) if want_dist else None
    ws_bytes = _lib.lib().pu3_group_knn_workspace(B, C, M, N, k, p_div, int(bool(unique)))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    idx64 = idx if idx_dtype == torch.int64 else None
    idx32 = idx if idx_dtype == torch.int32 else None
    # stream goes last in the C signature, after (workspace, workspace_bytes)
    _lib.launch("pu3_group_knn_f32", q, B, C, M, N, k, p_div, _lib.ptr(q), _lib.ptr(p), uflag,
                int(max_group or B), _lib.ptr(knn), _lib.ptr(idx64), _lib.ptr(idx32), _lib.ptr(dist), _lib.ptr(ws),
                ws_bytes, extra_kernels=_dup_prepass_kernels(unique, C, N, k),
                tag=f"pu3_group_knn_f32[c={C},k={k},n<={N}]")
    return knn, idx, dist


class GroupKNNFunction(torch.autograd.Function):
    """Differentiable like the reference's composition: neighbours through torch.gather
    (operations.py:209-211) and distances through the distance matrix (:158-161, :207)."""

    @staticmethod
    def forward(ctx, k, query, points, unique, max_group):
        knn, idx, dist = _knn_raw(k, query, points, unique, max_group)
        ctx.save_for_backward(query, points, idx)
        ctx.mark_non_differentiable(idx)
        return knn, idx, dist

    @staticmethod
    def backward(ctx, g_knn, _g_idx, g_dist):
        query, points, idx = ctx.saved_tensors
        B, C, M = query.shape
        Bp, _, N = points.shape
        k = idx.size(2)
        p_div = B // Bp
        g_query = None
        g_scatter = g_knn.contiguous() if g_knn is not None else None
        if g_dist is not None:
            # d/dq D = 2 (q - p_j), d/dp_j D = -2 (q - p_j); the duplicate penalty is a constant
            exp_idx = idx.unsqueeze(1).expand(-1, C, -1, -1).reshape(B, C, M * k)
            src = points if p_div == 1 else points.repeat_interleave(p_div, dim=0)
            nb = torch.gather(src, 2, exp_idx).view(B, C, M, k)
            diff = 2.0 * g_dist.unsqueeze(1) * (query.unsqueeze(-1) - nb)
            g_query = diff.sum(dim=-1)
            g_scatter = -diff if g_scatter is None else g_scatter - diff
        g_points = None
        if g_scatter is not None and ctx.needs_input_grad[2]:
            g_points = torch.zeros_like(points)
            g_scatter = g_scatter.contiguous()
            _lib.launch("pu3_group_gather_bwd_f32", points, B, C, M, N, k, p_div, _lib.ptr(g_scatter), _lib.ptr(idx),
                        _lib.ptr(g_points))
        if not ctx.needs_input_grad[1]:
            g_query = None
        return None, g_query, g_points, None, None


def group_knn(k, query, points, unique=True, NCHW=True, max_group=None):
    """operations.py:165-216.
    :param k neighbourhood size; query BxCxM or BxMxC; points BxCxN or BxNxC
    :param unique  duplicated points (all coordinates equal to an earlier point) are pushed to the back
    :param NCHW    second dimension is the channel dimension
    :param max_group (extension) number of consecutive batch elements the duplicate penalty max(D)
                   is taken over; None = the whole batch, which is the reference's behaviour
    :return neighbor_points BxCxMxk (NCHW) or BxMxkxC, index_batch BxMxk int64, distance_batch BxMxk ascending
    `points` may have a batch that divides the query batch: element i then reads points[i // (B/Bp)]
    (what the reference obtains with expand(), upsampler.py:319-323).
    """
    _need_cuda(query, "group_knn"); _need_cuda(points, "group_knn")
    if query.dtype != torch.float32 or points.dtype != torch.float32:
        raise RuntimeError("group_knn: float32 tensors required")
    if NCHW:
        q, p = query.contiguous(), points.contiguous()
    else:
        q, p = query.transpose(2, 1).contiguous(), points.transpose(2, 1).contiguous()
    assert p.size(2) >= k, "points size must be greater or equal to k"
    knn, idx, dist = GroupKNNFunction.apply(int(k), q, p, bool(unique), max_group)
    if not NCHW:
        knn = knn.permute(0, 2, 3, 1)
    return knn, idx, dist


# ------------------------------------------------------------------------------------------------
# gather_points / furthest_point_sample
# ------------------------------------------------------------------------------------------------
class GatherFunction(torch.autograd.Function):
    """operations.py:219-263."""

    @staticmethod
    def forward(ctx, features, idx):
        _need_cuda(features, "gather_points")
        features = features.contiguous()
        idx = idx.contiguous().to(dtype=torch.int32)
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, npoint, dtype=features.dtype, device=features.device)
        sampling.gather_forward(B, C, N, npoint, features, idx, output)
        ctx.save_for_backward(idx)
        ctx.C, ctx.N = C, N
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, = ctx.saved_tensors
        B, npoint = idx.size()
        grad_features = torch.zeros(B, ctx.C, ctx.N, dtype=grad_out.dtype, device=grad_out.device)
        sampling.gather_backward(B, ctx.C, ctx.N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_points = GatherFunction.apply


FPS_ONCHIP_MAX = 16 * 512 * 32     # csrc/fps.cu: 16-CTA cluster x 512 threads x 32 points per thread


class FurthestPointSampling(torch.autograd.Function):
    """operations.py:269-297.  xyz (B,N,3) -> idx (B,npoint) int32, not differentiable."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        _need_cuda(xyz, "furthest_point_sample")
        B, N, _ = xyz.size()
        idx = torch.empty([B, npoint], dtype=torch.int32, device=xyz.device)
        # temp = None: "filled with 1e10, not written back" -- the reference allocates and fills a (B,N) buffer per
        # call (operations.py:291) that nobody reads afterwards.  Clouds beyond the on-chip capacity of a cluster
        # (FPS_ONCHIP_MAX points) keep their running distances in global memory and need the buffer for real.
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device) if N > FPS_ONCHIP_MAX else None
        _lib.launch("pu3_fps_f32", xyz, B, N, int(npoint), _lib.ptr(xyz), _lib.ptr(temp), _lib.ptr(idx))
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad_idx):
        return None, None


def furthest_point_sample(xyz, npoint, NCHW=True):
    """operations.py:303-323.
    :param xyz (B,3,N) or (B,N,3); npoint a constant
    :return idx (B,npoint) int32 and the sampled points (B,3,npoint) or (B,npoint,3)
    """
    assert xyz.dim() == 3, "input for furthest sampling must be a 3D-tensor, but xyz.size() is {}".format(xyz.size())
    if NCHW:
        xyz = xyz.transpose(2, 1).contiguous()
    assert xyz.size(2) == 3, "furthest sampling is implemented for 3D points"
    if xyz.dtype != torch.float32:
        raise RuntimeError("furthest_point_sample: float32 required")
    idx = FurthestPointSampling.apply(xyz.contiguous(), npoint)
    sampled_pc = gather_points(xyz.transpose(2, 1).contiguous(), idx)
    if not NCHW:
        sampled_pc = sampled_pc.transpose(2, 1).contiguous()
    return idx, sampled_pc


def furthest_point_sample_ragged(xyz, n_arr, m_arr, m_max):
    """Batched FPS over clouds of different sizes (batched eval): xyz (B,3,Nmax) with n_arr[i] valid leading
    points, m_arr[i] <= m_max samples wanted (None = m_max everywhere).  Returns idx (B,m_max) int32 (zeros past
    m_arr[i]) and the gathered points (B,3,m_max).  Every cloud gets what furthest_point_sample gives it alone."""
    _need_cuda(xyz, "furthest_point_sample_ragged")
    B, _, N = xyz.shape
    pts = xyz.transpose(2, 1).contiguous()
    idx = torch.zeros(B, m_max, dtype=torch.int32, device=xyz.device)
    _lib.launch("pu3_fps_ragged_f32", pts, B, N, int(m_max), _lib.ptr(n_arr), _lib.ptr(m_arr), _lib.ptr(pts), None,
                _lib.ptr(idx), tag="pu3_fps_f32")
    return idx, gather_points(xyz.contiguous(), idx)
