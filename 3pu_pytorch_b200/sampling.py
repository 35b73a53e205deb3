"""Drop-in for the reference's `sampling` extension module (sampling/sampling.cpp:83-89).

Same function names, argument order and ownership rules: the caller allocates every output,
the callee fills it in place and returns the same tensor.  Underneath: libpu3_b200 (sm_100a).
Differences, all on the safe side: work is enqueued on the CURRENT stream of the tensor's device
(the reference uses the legacy default stream with no device guard), failures raise RuntimeError
(the reference calls exit(-1)), and FPS is correct for b > 32.
"""
import torch

from . import _lib

_ELEM = {torch.float16: 2, torch.float32: 4, torch.float64: 8}
_BWD_DTYPE = {torch.float32: 0, torch.float64: 1, torch.float16: 2}


def furthest_sampling(b, n, m, input, temp, idx):
    """sampling.cpp:26-35.  input (b,n,3) f32, temp (b,n) f32 pre-filled 1e10, idx (b,m) i32 -> idx."""
    _lib.require_cuda(input, "input"); _lib.require_contiguous(input, "input")
    _lib.require_cuda(temp, "temp"); _lib.require_contiguous(temp, "temp")
    _lib.require_cuda(idx, "idx"); _lib.require_contiguous(idx, "idx")
    if input.dtype != torch.float32 or temp.dtype != torch.float32 or idx.dtype != torch.int32:
        raise RuntimeError("furthest_sampling expects float32 input/temp and int32 idx")
    if input.numel() != b * n * 3 or temp.numel() != b * n or idx.numel() != b * m:
        raise RuntimeError("furthest_sampling: tensor sizes do not match (b, n, m)")
    _lib.launch("pu3_fps_f32", input, b, n, m, _lib.ptr(input), _lib.ptr(temp), _lib.ptr(idx))
    return idx


def gather_forward(b, c, n, npoints, points, idx, out):
    """sampling.cpp:37-45.  points (b,c,n) f16/f32/f64, idx (b,npoints) i32, out (b,c,npoints) -> out."""
    _lib.require_cuda(points, "points_tensor"); _lib.require_contiguous(points, "points_tensor")
    _lib.require_cuda(idx, "idx_tensor"); _lib.require_contiguous(idx, "idx_tensor")
    _lib.require_cuda(out, "out_tensor"); _lib.require_contiguous(out, "out_tensor")
    if points.dtype not in _ELEM or out.dtype != points.dtype or idx.dtype != torch.int32:
        raise RuntimeError("gather_forward expects half/float/double points and int32 idx")
    if points.numel() != b * c * n or idx.numel() != b * npoints or out.numel() != b * c * npoints:
        raise RuntimeError("gather_forward: tensor sizes do not match (b, c, n, npoints)")
    _lib.launch("pu3_gather_fwd", points, b, c, n, npoints, _ELEM[points.dtype], _lib.ptr(points), _lib.ptr(idx),
                                             _lib.ptr(out))
    return out


def gather_backward(b, c, n, npoints, grad_out, idx, grad_points):
    """sampling.cpp:47-53.  Accumulates grad_out (b,c,npoints) into the zero-filled grad_points (b,c,n)."""
    for t, nm in ((grad_out, "grad_out_tensor"), (idx, "idx_tensor"), (grad_points, "grad_points_tensor")):
        _lib.require_cuda(t, nm); _lib.require_contiguous(t, nm)
    if grad_out.dtype not in _BWD_DTYPE or grad_points.dtype != grad_out.dtype or idx.dtype != torch.int32:
        raise RuntimeError("gather_backward expects half/float/double grads and int32 idx")
    if grad_out.numel() != b * c * npoints or idx.numel() != b * npoints or grad_points.numel() != b * c * n:
        raise RuntimeError("gather_backward: tensor sizes do not match (b, c, n, npoints)")
    _lib.launch("pu3_gather_bwd", grad_out, b, c, n, npoints, _BWD_DTYPE[grad_out.dtype], _lib.ptr(grad_out),
                                             _lib.ptr(idx), _lib.ptr(grad_points))
    return grad_points


def ball_query(query, xyz, radius, nsample):
    """sampling.cpp:59-81 (never called by the reference's Python).  query (b,m,3), xyz (b,n,3) -> idx (b,m,nsample) i32."""
    _lib.require_cuda(query, "query"); _lib.require_contiguous(query, "query")
    _lib.require_cuda(xyz, "xyz"); _lib.require_contiguous(xyz, "xyz")
    if query.dtype != torch.float32 or xyz.dtype != torch.float32:
        raise RuntimeError("ball_query expects float32")
    idx = torch.zeros(query.size(0), query.size(1), nsample, dtype=torch.int32, device=query.device)
    _lib.launch("pu3_ball_query_f32", query, xyz.size(0), xyz.size(1), query.size(1), float(radius), int(nsample),
                                                 _lib.ptr(query), _lib.ptr(xyz), _lib.ptr(idx))
    return idx
