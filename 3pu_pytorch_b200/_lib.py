"""ctypes binding of libpu3_b200.so (the C ABI declared in include/pu3_b200.h).

There is no fallback of any kind: if the shared library is missing, or a call
returns a non-zero status, a RuntimeError is raised.  The reference kills the
process on a failed launch (sampling/sampling_cuda.cu:56-60) or ignores the
status (network/model_loss.py:15); raising is the only behavioural change.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpu3_b200.so")
ABI_VERSION = 1

_c_int, _c_void_p, _c_size_t, _c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float
_c_ll = ctypes.c_longlong

# name -> (restype, argtypes); must list every symbol of include/pu3_b200.h (tests check this)
SIGNATURES = {
    "pu3_last_error": (ctypes.c_char_p, []),
    "pu3_version": (_c_int, []),
    "pu3_device_info": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "pu3_fps_f32": (_c_int, [_c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "pu3_fps_ragged_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 6),
    "pu3_gather_fwd": (_c_int, [_c_int] * 5 + [_c_void_p] * 4),
    "pu3_gather_bwd": (_c_int, [_c_int] * 5 + [_c_void_p] * 4),
    "pu3_ball_query_f32": (_c_int, [_c_int, _c_int, _c_int, _c_float, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "pu3_nmdist_fwd_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 7),
    "pu3_nmdist_bwd_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 9),
    "pu3_group_knn_workspace": (_c_size_t, [_c_int] * 7),
    "pu3_knn_set_grid": (None, [_c_int]),
    "pu3_knn_no_prefilter": (None, [_c_int]),
    "pu3_group_knn_f32": (_c_int, [_c_int] * 6 + [_c_void_p] * 2 + [_c_int] * 2 + [_c_void_p] * 5 + [_c_size_t, _c_void_p]),
    "pu3_group_knn_ragged_f32": (_c_int, [_c_int] * 7 + [_c_void_p] * 6 + [_c_int] + [_c_void_p] * 5 + [_c_size_t, _c_void_p]),
    "pu3_group_gather_bwd_f32": (_c_int, [_c_int] * 6 + [_c_void_p] * 4),
    "pu3_pointwise_conv_f32": (_c_int, [_c_int] * 4 + [_c_void_p, _c_ll, _c_void_p, _c_void_p, _c_void_p, _c_ll,
                                                         _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_void_p]),
    "pu3_pointwise_conv_ex_f32": (_c_int, [_c_int] * 4 + [_c_void_p, _c_ll, _c_void_p, _c_void_p, _c_void_p, _c_ll,
                                                            _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_void_p, _c_ll, _c_int, _c_void_p]),
    "pu3_expand_code_f32": (_c_int, [_c_int] * 4 + [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "pu3_conv_tc_wsplit_bytes": (_c_size_t, [_c_int, _c_int]),
    "pu3_conv_tc_set_variant": (None, [_c_int]),
    "pu3_conv_tc_set_debug": (None, [_c_void_p]),
    "pu3_conv_tc_prepare_f32": (_c_int, [_c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "pu3_conv_tc_f32": (_c_int, [_c_int] * 4 + [_c_void_p, _c_ll, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_void_p]),
    "pu3_conv_tc_expand_f32": (_c_int, [_c_int] * 5 + [_c_void_p, _c_ll, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                                         _c_void_p, _c_void_p, _c_ll, _c_void_p]),
    "pu3_conv_tc_project_f32": (_c_int, [_c_int] * 5 + [_c_void_p, _c_ll, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                                          _c_void_p, _c_ll, _c_void_p, _c_ll, _c_int, _c_int, _c_void_p]),
    "pu3_head_tc_f32": (_c_int, [_c_int] * 3 + [_c_void_p, _c_ll] + [_c_void_p] * 4 + [_c_int, _c_int] + [_c_void_p] * 7
                        + [_c_ll, _c_void_p, _c_ll, _c_void_p]),
    "pu3_head_tc_set_debug": (None, [_c_void_p]),
    "pu3_head_tc_set_mode": (None, [_c_int]),
    "pu3_skip_fuse_f32": (_c_int, [_c_int] * 6 + [_c_void_p] * 7),
    "pu3_skip_force_generic": (None, [_c_int]),
    "pu3_edgeconv_set_tc": (None, [_c_int]),
    "pu3_to_point_major_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 4),
    "pu3_clip_adam_f32": (_c_int, [_c_ll] + [_c_void_p] * 4 + [_c_float] * 6 + [_c_int, _c_void_p]),
    "pu3_pointwise_conv_bwd_w_f32": (_c_int, [_c_int] * 4 + [_c_void_p, _c_ll, _c_void_p, _c_ll, _c_void_p, _c_void_p, _c_void_p]),
    "pu3_edgeconv_bwd_f32": (_c_int, [_c_int] * 3 + [_c_void_p, _c_ll, _c_void_p, _c_int, _c_int] + [_c_void_p] * 6 +
                             [_c_void_p, _c_ll, _c_void_p, _c_ll] + [_c_void_p] * 6 + [_c_void_p]),
    "pu3_level_workspace": (_c_size_t, [_c_int] * 8),
    "pu3_level_forward_f32": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                       _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "pu3_level_forward_pm_f32": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                          _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t,
                                          _c_void_p]),
    "pu3_level_forward_train_f32": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                             _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t,
                                             _c_void_p, _c_void_p]),
    "pu3_skip_fuse_ex_f32": (_c_int, [_c_int] * 6 + [_c_void_p] * 8),
    "pu3_skip_fuse_pm_f32": (_c_int, [_c_int] * 6 + [_c_void_p] * 9),
    "pu3_skip_bwd_f32": (_c_int, [_c_int] * 6 + [_c_void_p] * 6),
    "pu3_pointwise_conv_bwd_w_ex_f32": (_c_int, [_c_int] * 4 + [_c_void_p, _c_ll, _c_void_p, _c_ll, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "pu3_replica_sum_f32": (_c_int, [_c_ll, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p]),
    "pu3_relu_mask_f32": (_c_int, [_c_ll, _c_void_p, _c_void_p, _c_void_p]),
    "pu3_iota_i32": (_c_int, [_c_int, _c_void_p, _c_void_p]),
    "pu3_level_set_tc": (None, [_c_int]),
    "pu3_level_set_knn_override": (None, [_c_void_p] * 5),
    "pu3_normalize_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 5),
    "pu3_outlier_compact_f32": (_c_int, [_c_int] * 5 + [_c_void_p] * 10),
    "pu3_tile_seeds_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 5),
    "pu3_tiles_normalize_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 7),
    "pu3_denorm_merge_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 5),
    "pu3_gather_pm_f32": (_c_int, [_c_int] * 3 + [_c_void_p] * 4),
    "pu3_edgeconv_f32": (_c_int, [_c_int] * 3 + [_c_void_p, _c_ll, _c_void_p, _c_int, _c_int] + [_c_void_p] * 6 +
                         [_c_void_p, _c_ll, _c_void_p]),
    "pu3_edgeconv_ffma_f32": (_c_int, [_c_int] * 3 + [_c_void_p, _c_ll, _c_void_p, _c_int, _c_int] + [_c_void_p] * 6 +
                         [_c_void_p, _c_ll, _c_void_p]),
}

class LevelWeights(ctypes.Structure):
    """pu3_level_weights of include/pu3_b200.h"""
    _fields_ = [("layer0_w", _c_void_p), ("layer0_b", _c_void_p),
                ("ec_w", (_c_void_p * 3) * 4), ("ec_b", (_c_void_p * 3) * 4),
                ("prep_w", _c_void_p * 3), ("prep_b", _c_void_p * 3),
                ("up1_w", _c_void_p), ("up1_w_feat", _c_void_p), ("up1_b", _c_void_p),
                ("up2_w", _c_void_p), ("up2_b", _c_void_p), ("fc1_w", _c_void_p), ("fc1_b", _c_void_p),
                ("fc2_w", _c_void_p), ("fc2_b", _c_void_p), ("code", _c_void_p),
                ("r", _c_int), ("knn", _c_int), ("fm_knn", _c_int), ("reserved", _c_int)]


class LevelSaved(ctypes.Structure):
    """pu3_level_saved of include/pu3_b200.h"""
    _fields_ = [("h", _c_void_p * 4), ("idx", _c_void_p * 4), ("skip_idx", _c_void_p), ("skip_w", _c_void_p),
                ("h1", _c_void_p), ("h2", _c_void_p), ("feat_pre", _c_void_p)]


_lib = None

# Bumped by every writer that changes parameters through raw pointers (dist.FlatAdam.step): tensor._version does not
# see those writes, so caches of weight-derived device images (upsampler.Level._engine_weights) key on this as well.
weight_generation = 0


def bump_weight_generation():
    global weight_generation
    weight_generation += 1


def lib():
    """The loaded library; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"libpu3_b200.so not found at {LIB_PATH}: build it with `make -C {_HERE}/csrc` "
                "(or __graft_entry__.build()); there is no CPU or PyTorch fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype, fn.argtypes = res, args
        if handle.pu3_version() != ABI_VERSION:
            raise RuntimeError(f"libpu3_b200.so ABI {handle.pu3_version()} != expected {ABI_VERSION}: rebuild")
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().pu3_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


# kernels enqueued per entry point (for the launch count bench.py reports); memsets are not kernels
KERNELS_PER_CALL = {
    "pu3_fps_f32": 1, "pu3_gather_fwd": 1, "pu3_gather_bwd": 1, "pu3_ball_query_f32": 1, "pu3_nmdist_fwd_f32": 1,
    "pu3_nmdist_bwd_f32": 1, "pu3_group_gather_bwd_f32": 1, "pu3_pointwise_conv_f32": 1, "pu3_expand_code_f32": 1,
    "pu3_edgeconv_f32": 1, "pu3_edgeconv_ffma_f32": 1, "pu3_iota_i32": 1,
    # layer0 + 4 x (kNN + edge-conv) + 3 x (prep weight split + prep conv) + 4 head kernels (3 weight splits + 1 fused tcgen05; the
    # train-mode forward keeps the three-kernel head: 21);
    # the feature kNN finds duplicates itself (no side kernels); the skip connection adds 3 duplicate kernels + kNN + skip, iota 1
    "pu3_level_forward_f32": 19, "pu3_level_forward_pm_f32": 19, "pu3_level_forward_train_f32": 21,
    "pu3_conv_tc_prepare_f32": 1, "pu3_conv_tc_f32": 1, "pu3_conv_tc_expand_f32": 1, "pu3_conv_tc_project_f32": 1,
    "pu3_head_tc_f32": 1,
    "pu3_fps_ragged_f32": 1,
    "pu3_group_knn_f32": 1, "pu3_group_knn_ragged_f32": 1,  # + 3 (duplicate flags, group flags, max D) when unique
}


class Profiler:
    """Per-entry-point CUDA-event timing on the launching stream + launch counter (bench.py, tests).
    Events are recorded only while a Profiler is installed; the cost is two cudaEventRecord per call."""

    def __init__(self, timing=True):
        self.timing = timing
        self.launches = 0
        self.calls = {}
        self._events = {}

    def add(self, name, nkernels, e0, e1):
        self.launches += nkernels
        self.calls[name] = self.calls.get(name, 0) + 1
        if e0 is not None:
            self._events.setdefault(name, []).append((e0, e1))

    ENGINE_TAGS = ["pu3_pointwise_conv_f32", "pu3_group_knn_f32[c=24,k=33,n<=312]", "pu3_edgeconv_f32",
                   "pu3_group_knn_f32[c=3,k=5,skip]", "pu3_skip_fuse_f32", "pu3_expand_code_f32", "misc",
                   "pu3_head_tc_f32[tcgen05 head: 3 weight splits + ONE fused up1/up2/fc1/fc2 kernel]",
                   "pu3_conv_tc_f32[tcgen05 prep convs 84/144/204->24: 3 weight splits + 3 convs]"]

    def summary(self):
        """{name: (calls, total_ms)} -- call after torch.cuda.synchronize().  Kernels launched inside the level
        engine are reported under their own families (timed by the engine's event pairs), not under the engine."""
        out = {}
        for name, n in self.calls.items():
            ms = sum(a.elapsed_time(b) for a, b in self._events.get(name, []))
            out[name] = (n, ms)
        if self.timing:
            n = len(self.ENGINE_TAGS)
            ms = (ctypes.c_float * n)(); calls = (ctypes.c_int * n)()
            h = ctypes.CDLL(LIB_PATH)
            h.pu3_prof_collect(ms, calls, n)
            inner = 0.0
            for i, name in enumerate(self.ENGINE_TAGS):
                if calls[i]:
                    c0, m0 = out.get(name, (0, 0.0))
                    out[name] = (c0 + calls[i], m0 + ms[i]); inner += ms[i]
            if "pu3_level_forward_f32" in out and inner > 0:
                c0, m0 = out.pop("pu3_level_forward_f32")
                out["pu3_level_forward_f32[engine: gaps between its kernels]"] = (c0, max(m0 - inner, 0.0))
        return out


_profiler = None


def set_profiler(p):
    global _profiler
    prev, _profiler = _profiler, p
    ctypes.CDLL(LIB_PATH).pu3_prof_enable(1 if (p is not None and p.timing) else 0)
    return prev


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)     # (device index) -> cudaStream_t as int, no Stream object
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def launch(name, tensor, *args, extra_kernels=0, tag=None):
    """Call entry point `name` with `args` (+ the current stream of `tensor`'s device appended), on that device,
    and raise on a non-zero status.  The common case (no profiler, tensor on the current device) is kept short: this is the
    per-launch host cost of every eager call (a 5 us kernel was taking 16 us to issue through the torch.cuda wrappers)."""
    fn = getattr(_lib or lib(), name)
    prof = _profiler
    idx = tensor.device.index
    if prof is None and _raw_stream is not None and (idx is None or idx == _raw_device()):
        status = fn(*args, _raw_stream(idx if idx is not None else _raw_device()))
        if status != 0:
            check(status, name)
        return
    with on_device(tensor):
        stream = torch.cuda.current_stream(tensor.device)
        if prof is not None and prof.timing:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            status = fn(*args, stream.cuda_stream)
            e1.record(stream)
        else:
            e0 = e1 = None
            status = fn(*args, stream.cuda_stream)
    if status != 0:
        check(status, name)
    if prof is not None:
        prof.add(tag or name, KERNELS_PER_CALL.get(name, 1) + extra_kernels, e0, e1)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(t, name):
    # same wording as the reference's CHECK_CUDA / CHECK_CONTIGUOUS (sampling/sampling.cpp:20-24)
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def require_contiguous(t, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


class on_device:
    """Make the tensor's device current for the duration of a call (the reference has no device guard)."""

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and self.idx != cur:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False
