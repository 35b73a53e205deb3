"""ctypes front-end of oracle_c.c (CPU restatement of the reference's CUDA extensions).

TEST INFRASTRUCTURE ONLY -- see the header of oracle_c.c.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.

All functions take / return numpy arrays (C-contiguous) and mirror the argument
meaning of the reference entry points:
  sampling.furthest_sampling / gather_forward / gather_backward   (sampling/sampling.cpp:26-53)
  losses.nmdistance_forward / nmdistance_backward                 (losses/nmdistance.cpp:12-27)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def build(force=False):
    """Compile oracle_c.c with gcc (a second or two)."""
    src = os.path.join(_HERE, "oracle_c.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_c.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def opt_n_threads(n):
    return int(lib().oracle_opt_n_threads(ctypes.c_int(int(n))))


def fps(xyz, m, temp=None, legacy_temp_rows=False):
    """xyz (b,n,3) f32 -> idx (b,m) i32.  temp (b,n) in/out, default 1e10 (operations.py:291)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    b, n, _ = xyz.shape
    if temp is None:
        temp = np.full((b, n), 1e10, dtype=np.float32)
    assert temp.dtype == np.float32 and temp.flags.c_contiguous
    idx = np.zeros((b, m), dtype=np.int32)
    lib().oracle_fps(b, n, m, _p(xyz), _p(temp), _p(idx), int(bool(legacy_temp_rows)))
    return idx


def gather_fwd(points, idx):
    """points (b,c,n) f16/f32/f64, idx (b,m) i32 -> (b,c,m)."""
    points = np.ascontiguousarray(points)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), dtype=points.dtype)
    fn = {np.dtype(np.float32): "oracle_gather_fwd_f32", np.dtype(np.float64): "oracle_gather_fwd_f64",
          np.dtype(np.float16): "oracle_gather_fwd_u16"}[points.dtype]
    getattr(lib(), fn)(b, c, n, m, _p(points), _p(idx), _p(out))
    return out


def gather_bwd(grad_out, idx, n, grad_points=None):
    """grad_out (b,c,m), idx (b,m) -> grad_points (b,c,n) accumulated (zero-initialised if None)."""
    grad_out = np.ascontiguousarray(grad_out)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    b, c, m = grad_out.shape
    if grad_points is None:
        grad_points = np.zeros((b, c, n), dtype=grad_out.dtype)
    fn = {np.dtype(np.float32): "oracle_gather_bwd_f32", np.dtype(np.float64): "oracle_gather_bwd_f64"}[grad_out.dtype]
    getattr(lib(), fn)(b, c, n, m, _p(grad_out), _p(idx), _p(grad_points))
    return grad_points


def nmdist_fwd(xyz1, xyz2):
    """xyz1 (b,n,3), xyz2 (b,m,3) f32 -> dist1 (b,n), idx1 (b,n), dist2 (b,m), idx2 (b,m)."""
    xyz1 = np.ascontiguousarray(xyz1, dtype=np.float32)
    xyz2 = np.ascontiguousarray(xyz2, dtype=np.float32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((b, n), np.float32)
    i1 = np.zeros((b, n), np.int32)
    d2 = np.zeros((b, m), np.float32)
    i2 = np.zeros((b, m), np.int32)
    lib().oracle_nmdist_fwd(b, n, m, _p(xyz1), _p(xyz2), _p(d1), _p(i1), _p(d2), _p(i2))
    return d1, i1, d2, i2


def nmdist_bwd(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """-> gradxyz1 (b,n,3), gradxyz2 (b,m,3), zero-initialised as model_loss.py:25-26 does."""
    xyz1 = np.ascontiguousarray(xyz1, dtype=np.float32)
    xyz2 = np.ascontiguousarray(xyz2, dtype=np.float32)
    g1 = np.ascontiguousarray(graddist1, dtype=np.float32)
    g2 = np.ascontiguousarray(graddist2, dtype=np.float32)
    i1 = np.ascontiguousarray(idx1, dtype=np.int32)
    i2 = np.ascontiguousarray(idx2, dtype=np.int32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.zeros_like(xyz1)
    gx2 = np.zeros_like(xyz2)
    lib().oracle_nmdist_bwd(b, n, m, _p(xyz1), _p(xyz2), _p(g1), _p(g2), _p(i1), _p(i2), _p(gx1), _p(gx2))
    return gx1, gx2


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz = np.ascontiguousarray(new_xyz, dtype=np.float32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    lib().oracle_ball_query(b, n, m, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx))
    return idx
