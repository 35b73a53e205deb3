"""Drop-in for the reference's `losses` extension module (losses/nmdistance.cpp:24-27).

nmdistance_forward / nmdistance_backward keep the reference's argument order, caller-allocated
outputs and the int return (1 = ok); a failed launch raises instead of printing and returning 0.
"""
import torch

from . import _lib


def _check3(t, name):
    _lib.require_cuda(t, name); _lib.require_contiguous(t, name)
    if t.dtype != torch.float32 or t.dim() != 3 or t.size(2) != 3:
        raise RuntimeError(f"{name} must be a float32 (B,N,3) tensor")


def nmdistance_forward(xyz1, xyz2, dist1, dist2, idx1, idx2):
    """nmdistance.cpp:12-14.  xyz1 (B,N,3), xyz2 (B,M,3) -> dist1/idx1 (B,N), dist2/idx2 (B,M) filled in place."""
    _check3(xyz1, "xyz1"); _check3(xyz2, "xyz2")
    b, n, m = xyz1.size(0), xyz1.size(1), xyz2.size(1)
    for t, nm, dt, cnt in ((dist1, "dist1", torch.float32, b * n), (dist2, "dist2", torch.float32, b * m),
                           (idx1, "idx1", torch.int32, b * n), (idx2, "idx2", torch.int32, b * m)):
        _lib.require_cuda(t, nm); _lib.require_contiguous(t, nm)
        if t.dtype != dt or t.numel() != cnt:
            raise RuntimeError(f"nmdistance_forward: {nm} has the wrong dtype or size")
    if xyz2.size(0) != b:
        raise RuntimeError("nmdistance_forward: batch sizes differ")
    _lib.launch("pu3_nmdist_fwd_f32", xyz1, b, n, m, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(dist1), _lib.ptr(dist2),
                                                 _lib.ptr(idx1), _lib.ptr(idx2))
    return 1


def nmdistance_backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
    """nmdistance.cpp:17-21.  Accumulates into the zero-filled gradxyz1 (B,N,3), gradxyz2 (B,M,3)."""
    _check3(xyz1, "xyz1"); _check3(xyz2, "xyz2"); _check3(gradxyz1, "gradxyz1"); _check3(gradxyz2, "gradxyz2")
    b, n, m = xyz1.size(0), xyz1.size(1), xyz2.size(1)
    # autograd hands over expand()ed (stride-0) gradients after a mean(); the reference reads them
    # through .data<float>() as if they were dense (nmdistance_cuda.cu:183) -- densify instead
    graddist1, graddist2 = graddist1.contiguous(), graddist2.contiguous()
    for t, nm, dt, cnt in ((graddist1, "graddist1", torch.float32, b * n), (graddist2, "graddist2", torch.float32, b * m),
                           (idx1, "idx1", torch.int32, b * n), (idx2, "idx2", torch.int32, b * m)):
        _lib.require_cuda(t, nm); _lib.require_contiguous(t, nm)
        if t.dtype != dt or t.numel() != cnt:
            raise RuntimeError(f"nmdistance_backward: {nm} has the wrong dtype or size")
    _lib.launch("pu3_nmdist_bwd_f32", xyz1, b, n, m, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(gradxyz1),
                                                 _lib.ptr(gradxyz2), _lib.ptr(graddist1), _lib.ptr(graddist2),
                                                 _lib.ptr(idx1), _lib.ptr(idx2))
    return 1
