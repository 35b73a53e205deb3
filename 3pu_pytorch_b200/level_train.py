"""Train-mode Level (network/upsampler.py:272-374 under autograd, model.py:53-66) as ONE autograd node.

forward  = the level engine (csrc/level.cu: layer0, 4 x {prep, feature kNN, fused DenseEdgeConv}, skip connection, tcgen05
           expansion head) keeping what the backward needs (pu3_level_saved);
backward = the chain rule of the same graph on libpu3_b200 kernels, in reverse:
           head (4 weight-gradient GEMMs + 4 input-gradient GEMMs with the ReLU mask in their epilogue, replica sums for
           the r-fold point replication and the residual), skip connection (row scatter into the previous level's
           point-major gradient; the weights are detached in the reference, upsampler.py:243,249), the four dense blocks
           (pu3_edgeconv_bwd_f32 + prep convolution) and layer0.
What autograd sees: (xyz_normalized, previous features, 40 parameters) -> (new xyz, point features).  The neighbour
searches carry no gradient (indices), like in the reference.
"""
import ctypes

import torch

from . import _lib

# FlatAdam (dist.py) points every parameter's .grad into one pre-zeroed flat buffer.  When set, the backward kernels
# accumulate straight into those views (they are "+=" kernels) and autograd receives None for the parameters: no
# per-parameter temporaries, no 160 AccumulateGrad additions per step.
accumulate_into_param_grads = False


def _t(w):
    """(cout, cin[,1[,1]]) weight -> contiguous (cin, cout) transpose for the input-gradient GEMM."""
    return w.reshape(w.shape[0], w.shape[1]).t().contiguous()


def _conv(x, xs, w, y, ys, b, n, cin, cout, bias=None, relu=False, mask=None, ms=0, acc=False):
    """y (pointer, batch stride ys) [+]= mask(act(w x + bias)); x a tensor or a raw pointer into one (w gives the device)."""
    _lib.launch("pu3_pointwise_conv_ex_f32", w, b, n, cin, cout, x.data_ptr() if torch.is_tensor(x) else x, xs, w.data_ptr(),
                _lib.ptr(bias), y, ys, None, 0, 1, 1, int(relu), mask, ms, int(acc))


class LevelTrainFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, level, xyz, xyz_norm, prev_xyz, prev_feat, max_group, *params):
        dev = xyz_norm.device
        T, _, N = xyz_norm.shape
        W = level._engine_weights(dev)
        r, K, fm = W.r, W.knn, W.fm_knn
        f32 = dict(dtype=torch.float32, device=dev)
        xn = xyz_norm.contiguous()
        has_prev = prev_feat is not None and fm > 0
        if has_prev:
            prev_xyz_c = prev_xyz.contiguous()
            Bp, Cp, No = prev_feat.shape
            if T % Bp != 0:
                raise RuntimeError(f"Level: previous batch {Bp} must divide batch {T}")
            prev_pm = torch.empty(Bp, No, Cp, **f32)                      # point-major rows for the gathers
            pf = prev_feat.contiguous()
            _lib.launch("pu3_to_point_major_f32", pf, Bp, Cp, No, pf.data_ptr(), None, prev_pm.data_ptr())
            xyz_c = xyz.contiguous()
        else:
            prev_xyz_c = prev_pm = xyz_c = None
            Bp = No = 0
        sv = _lib.LevelSaved()
        hs = [torch.empty(T, 24, N, **f32) for _ in range(4)]
        idxs = [torch.empty(T, N, K + 1, dtype=torch.int32, device=dev) for _ in range(4)]
        h1 = torch.empty(T, 128, N * r, **f32)
        h2 = torch.empty(T, 128, N * r, **f32)
        skip_idx = torch.empty(T, N, fm, dtype=torch.int64, device=dev) if has_prev else None
        skip_w = torch.empty(T, N, fm, **f32) if has_prev else None
        feat_pre = torch.empty(T, 264, N, **f32) if has_prev else None      # the prep convolutions read the features before the skip update
        for i in range(4):
            sv.h[i], sv.idx[i] = hs[i].data_ptr(), idxs[i].data_ptr()
        sv.skip_idx, sv.skip_w, sv.h1, sv.h2 = _lib.ptr(skip_idx), _lib.ptr(skip_w), h1.data_ptr(), h2.data_ptr()
        sv.feat_pre = _lib.ptr(feat_pre)
        L = _lib.lib()
        ws_bytes = L.pu3_level_workspace(T, N, r, K, fm, max(Bp, 1), max(No, 1), int(has_prev))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        feat = torch.empty(T, 264, N, **f32)
        out = torch.empty(T, 3, N * r, **f32)
        _lib.launch("pu3_level_forward_train_f32", xn, ctypes.addressof(W), T, N, _lib.ptr(xyz_c), xn.data_ptr(), None, 0,
                    int(max_group or T), _lib.ptr(prev_xyz_c), _lib.ptr(prev_pm), int(Bp), int(No), None, feat.data_ptr(),
                    out.data_ptr(), ws.data_ptr(), ws_bytes, ctypes.addressof(sv),
                    extra_kernels=5 if has_prev else 0)
        ctx.level, ctx.has_prev, ctx.dims = level, has_prev, (T, N, r, K, fm, Bp, No)
        ctx.save_for_backward(xn, feat, h1, h2, skip_idx, skip_w, feat_pre, *hs, *idxs, *params)
        return out, feat

    @staticmethod
    def backward(ctx, d_out, d_feat):
        level = ctx.level
        T, N, r, K, fm, Bp, No = ctx.dims
        saved = ctx.saved_tensors
        xn, feat, h1, h2, skip_idx, skip_w, feat_pre = saved[:7]
        hs, idxs, params = saved[7:11], saved[11:15], saved[15:]
        if feat_pre is None:
            feat_pre = feat
        dev = xn.device
        f32 = dict(dtype=torch.float32, device=dev)
        C, Nr = 264, N * r
        it = iter(params)
        l0_w, l0_b = next(it), next(it)
        ec = [[(next(it), next(it)) for _ in range(3)] for _ in range(4)]
        prep = [(next(it), next(it)) for _ in range(3)]
        up1_w, up1_b, up2_w, up2_b, fc1_w, fc1_b, fc2_w, fc2_b = (next(it) for _ in range(8))

        direct = accumulate_into_param_grads and all(p.grad is not None and p.grad.is_contiguous() for p in params)
        grads = {}

        def gbuf(p):
            """gradient accumulator of parameter p, viewed as (rows, cols)"""
            if direct:
                g = p.grad
            else:
                g = grads.get(id(p))
                if g is None:
                    g = grads[id(p)] = torch.zeros_like(p)
            return g

        def dW(x, xs, dy, dys, n, cin, cout, w, b, dw_ptr=None, dw_stride=None):
            gw = gbuf(w)
            _lib.launch("pu3_pointwise_conv_bwd_w_ex_f32", xn, T, n, cin, cout, x, xs, dy, dys,
                        dw_ptr if dw_ptr is not None else gw.data_ptr(), dw_stride or cin, gbuf(b).data_ptr() if b is not None else None)

        if d_out is None:
            d_out = torch.zeros(T, 3, Nr, **f32)
        d_out = d_out.contiguous()
        # ---- expansion head (:349-372), in reverse ------------------------------------------------------------------------
        dxn = torch.empty(T, 3, N, **f32)                                    # residual: xyz_normalized replicated r times (:371)
        _lib.launch("pu3_replica_sum_f32", xn, T * 3, N, r, d_out.data_ptr(), dxn.data_ptr(), 0)
        h3 = torch.empty(T, 64, Nr, **f32)                                   # fc_layer1 output: recomputed (the forward keeps it on chip)
        _conv(h2, 128 * Nr, fc1_w, h3.data_ptr(), 64 * Nr, T, Nr, 128, 64, bias=fc1_b, relu=True)
        dW(h3.data_ptr(), 64 * Nr, d_out.data_ptr(), 3 * Nr, Nr, 64, 3, fc2_w, fc2_b)
        dh3 = torch.empty(T, 64, Nr, **f32)
        _conv(d_out, 3 * Nr, _t(fc2_w), dh3.data_ptr(), 64 * Nr, T, Nr, 3, 64, mask=h3.data_ptr(), ms=64 * Nr)
        dW(h2.data_ptr(), 128 * Nr, dh3.data_ptr(), 64 * Nr, Nr, 128, 64, fc1_w, fc1_b)
        dh2 = torch.empty(T, 128, Nr, **f32)
        _conv(dh3, 64 * Nr, _t(fc1_w), dh2.data_ptr(), 128 * Nr, T, Nr, 64, 128, mask=h2.data_ptr(), ms=128 * Nr)
        dW(h1.data_ptr(), 128 * Nr, dh2.data_ptr(), 128 * Nr, Nr, 128, 128, up2_w, up2_b)
        dh1 = torch.empty(T, 128, Nr, **f32)
        _conv(dh2, 128 * Nr, _t(up2_w), dh1.data_ptr(), 128 * Nr, T, Nr, 128, 128, mask=h1.data_ptr(), ms=128 * Nr)
        # up_layer1 on [features replicated r times ; code]: the replicas share the feature columns of W
        dh1s = torch.empty(T, 128, N, **f32)
        _lib.launch("pu3_replica_sum_f32", xn, T * 128, N, r, dh1.data_ptr(), dh1s.data_ptr(), 0)
        g_up1 = gbuf(up1_w)
        dW(feat.data_ptr(), C * N, dh1s.data_ptr(), 128 * N, N, C, 128, up1_w, up1_b, dw_ptr=g_up1.data_ptr(), dw_stride=C + 1)
        code_row = level._code_row(dev, T, N)                                 # (T,1,N*r): code[p % r]
        dW(code_row.data_ptr(), Nr, dh1.data_ptr(), 128 * Nr, Nr, 1, 128, up1_w, None,
           dw_ptr=g_up1.data_ptr() + 4 * C, dw_stride=C + 1)
        w_up1_t = up1_w.reshape(128, C + 1)[:, :C].t().contiguous()           # (264,128)
        if d_feat is not None:
            dfeat = d_feat.contiguous().clone()
            _conv(dh1s, 128 * N, w_up1_t, dfeat.data_ptr(), C * N, T, N, 128, C, acc=True)
        else:
            dfeat = torch.empty(T, C, N, **f32)
            _conv(dh1s, 128 * N, w_up1_t, dfeat.data_ptr(), C * N, T, N, 128, C)
        # ---- inter-level skip connection (:317-347): gradient to the previous level's features ----------------------------------
        d_prev = None
        if ctx.has_prev and ctx.needs_input_grad[4]:
            dprev_pm = torch.zeros(Bp, No, C, **f32)
            _lib.launch("pu3_skip_bwd_f32", xn, T, N, C, fm, T // Bp, No, dfeat.data_ptr(), skip_idx.data_ptr(), skip_w.data_ptr(),
                        None, dprev_pm.data_ptr())
            d_prev = torch.empty(Bp, C, No, **f32)                            # back to channel-major: a (No, C) -> (C, No) transpose
            _lib.launch("pu3_to_point_major_f32", xn, Bp, No, C, dprev_pm.data_ptr(), None, d_prev.data_ptr())
        # ---- dense blocks, last to first (:288-311) -------------------------------------------------------------------------------
        fs = C * N
        for blk in (3, 2, 1, 0):
            s = 240 - 60 * (blk + 1)                                          # this block's 60 output channels: feat[:, s:s+60]
            (w0, b0), (w1, b1), (w2, b2) = ec[blk]
            if blk > 0:
                dh = torch.zeros(T, 24, N, **f32)
                dx_ptr, dx_stride = dh.data_ptr(), 24 * N
            else:                                                             # block 0 reads layer0's output = feat[:, 240:264]
                dh = None
                dx_ptr, dx_stride = dfeat.data_ptr() + 4 * 240 * N, fs
            gws = [gbuf(p) for p in (w0, b0, w1, b1, w2, b2)]
            _lib.launch("pu3_edgeconv_bwd_f32", xn, T, N, K, hs[blk].data_ptr(), 24 * N, idxs[blk].data_ptr(), K + 1, 1,
                        w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                        dfeat.data_ptr() + 4 * s * N, fs, dx_ptr, dx_stride, *[g.data_ptr() for g in gws])
            if blk > 0:
                pw, pb = prep[blk - 1]
                cin = C - (s + 60)
                _lib.launch("pu3_relu_mask_f32", xn, T * 24 * N, dh.data_ptr(), hs[blk].data_ptr())
                src = feat_pre.data_ptr() + 4 * (s + 60) * N
                dW(src, fs, dh.data_ptr(), 24 * N, N, cin, 24, pw, pb)
                _conv(dh, 24 * N, _t(pw), dfeat.data_ptr() + 4 * (s + 60) * N, fs, T, N, 24, cin, acc=True)
        # ---- layer0 (:288) -------------------------------------------------------------------------------------------------------
        dx0 = dfeat.data_ptr() + 4 * 240 * N
        dW(xn.data_ptr(), 3 * N, dx0, fs, N, 3, 24, l0_w, l0_b)
        d_xn = None
        if ctx.needs_input_grad[2]:
            _conv(dx0, fs, _t(l0_w), dxn.data_ptr(), 3 * N, T, N, 24, 3, acc=True)
            d_xn = dxn
        if direct:
            pgrads = [None] * len(params)
        else:
            pgrads = [grads.get(id(p)) for p in params]
        return (None, None, d_xn, None, d_prev, None, *pgrads)
