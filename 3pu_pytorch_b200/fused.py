"""Python front-ends of the fused compute kernels (1x1 convolutions, DenseEdgeConv, expansion head).

Forward passes run on libpu3_b200.  Gradients: when autograd is recording and an input or weight
requires grad, the same mathematics is evaluated through differentiable operators (group_knn's own
backward + torch autograd for the small dense algebra) so that training is correct today; the
hand-written backward kernels replace that branch kernel by kernel ("backward").
"""
import torch
import torch.nn.functional as F

from . import _lib
from .operations import group_knn, _knn_raw


native_edgeconv_backward = True   # False: differentiate DenseEdgeConv through the operator composition (tests)


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _check_f32_cuda(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError(f"{name}: float32 CUDA tensor required (3pu_pytorch_b200 has no CPU fallback)")


def conv1x1_autograd(x, weight, bias):
    """Differentiable 1x1 convolution as an fp32 matmul.  NOT F.conv2d: cuDNN convolutions run with TF32
    inputs by default (torch.backends.cudnn.allow_tf32), whose 5e-4 relative rounding breaks the 1e-5
    tolerance of this path; torch.matmul stays in full fp32 unless the user opts in."""
    shape = x.shape
    w2 = weight.reshape(weight.shape[0], weight.shape[1])
    y = torch.matmul(w2, x.reshape(shape[0], shape[1], -1))
    if bias is not None:
        y = y + bias.view(1, -1, 1)
    return y.reshape(shape[0], weight.shape[0], *shape[2:])


def conv_into(x, w, b, out, relu=False, residual=None, res_div=1):
    """out[:, :, :] = act(W x + b) (+ residual[..., p // res_div]) with x (B,Cin,N) and out (B,Cout,N) possibly
    channel slices of larger contiguous buffers (stride(1) == N, stride(2) == 1)."""
    B, Cin, N = x.shape
    Cout = out.shape[1]
    assert x.stride(2) == 1 and x.stride(1) == N and out.stride(2) == 1 and out.stride(1) == N
    w2 = w.reshape(Cout, Cin)
    if not w2.is_contiguous():
        w2 = w2.contiguous()
    rp, rbs, rn = None, 0, 1
    if residual is not None:
        assert residual.is_contiguous() and residual.shape[0] == B and residual.shape[1] == Cout
        rp, rbs, rn = residual.data_ptr(), residual.stride(0), residual.shape[2]
    _lib.launch("pu3_pointwise_conv_f32", x, B, N, Cin, Cout, x.data_ptr(), x.stride(0) if B > 1 else Cin * N,
                                                     w2.data_ptr(), _lib.ptr(b), out.data_ptr(),
                                                     out.stride(0) if B > 1 else Cout * N, rp, rbs, rn, res_div,
                                                     int(relu))
    return out


class PointwiseConvFunction(torch.autograd.Function):
    """1x1 convolution (+ReLU) with hand-written forward AND backward kernels: forward / input gradient on the
    FFMA2 SGEMM (csrc/pointwise_conv.cu), weight and bias gradients on its split-K companion."""

    @staticmethod
    def forward(ctx, x3, w2, bias, relu):
        B, Cin, N = x3.shape
        out = torch.empty(B, w2.shape[0], N, dtype=torch.float32, device=x3.device)
        conv_into(x3, w2, bias, out, relu=relu)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x3, w2, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x3, w2, out = ctx.saved_tensors
        B, Cin, N = x3.shape
        Cout = w2.shape[0]
        dy = dy.contiguous()
        if ctx.relu:
            dy = dy * (out > 0).to(dy.dtype)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x3)
            conv_into(dy, w2.t().contiguous(), None, dx)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.zeros_like(w2)
            db = torch.zeros(Cout, dtype=torch.float32, device=dy.device) if ctx.has_bias else None
            _lib.launch("pu3_pointwise_conv_bwd_w_f32", x3, B, N, Cin, Cout, x3.data_ptr(), Cin * N, dy.data_ptr(), Cout * N,
                        dw.data_ptr(), _lib.ptr(db))
        return dx, dw, db, None


def pointwise_conv(x, weight, bias, relu=False):
    """1x1 Conv1d/Conv2d (+ReLU) on (B,C,N) or (B,C,N,1) input: layers.py:115-204 with kernel size 1."""
    _check_f32_cuda(x, "pointwise_conv")
    if _needs_grad(x, weight, bias):
        shape = x.shape
        x3 = x.reshape(shape[0], shape[1], -1).contiguous()
        w2 = weight.reshape(weight.shape[0], weight.shape[1])
        y = PointwiseConvFunction.apply(x3, w2, bias, bool(relu))
        return y.reshape(shape[0], weight.shape[0], *shape[2:])
    shape = x.shape
    x3 = x.reshape(shape[0], shape[1], -1).contiguous()
    out = torch.empty(shape[0], weight.shape[0], x3.shape[2], dtype=torch.float32, device=x.device)
    conv_into(x3, weight, bias, out, relu=relu)
    return out.reshape(shape[0], weight.shape[0], *shape[2:])


def edgeconv_into(x, idx32, idx_off, k, weights, biases, out, ffma=False):
    """out (B,60,N) slice <- fused DenseEdgeConv of x (B,24,N) slice with neighbours idx32[..., idx_off:idx_off+k].
    ffma=True: the FFMA kernels only (the arithmetic the backward kernel recomputes) -- for forwards that will be differentiated."""
    B, C, N = x.shape
    assert C == 24 and out.shape[1] == 60 and idx32.dtype == torch.int32 and idx32.is_contiguous()
    assert x.stride(2) == 1 and x.stride(1) == N and out.stride(2) == 1 and out.stride(1) == N
    w = [wi.reshape(wi.shape[0], wi.shape[1]).contiguous() for wi in weights]
    _lib.launch("pu3_edgeconv_ffma_f32" if ffma else "pu3_edgeconv_f32", x, B, N, k, x.data_ptr(), x.stride(0) if B > 1 else C * N, idx32.data_ptr(),
                                               idx32.shape[2], idx_off, w[0].data_ptr(), biases[0].data_ptr(),
                                               w[1].data_ptr(), biases[1].data_ptr(), w[2].data_ptr(),
                                               biases[2].data_ptr(), out.data_ptr(),
                                               out.stride(0) if B > 1 else 60 * N)
    return out


def _edgeconv_supported(x, weights):
    return (x.shape[1] == 24 and len(weights) == 3 and tuple(weights[0].shape[:2]) == (12, 48)
            and tuple(weights[1].shape[:2]) == (12, 36) and tuple(weights[2].shape[:2]) == (12, 48))


class DenseEdgeConvFunction(torch.autograd.Function):
    """Fused DenseEdgeConv with hand-written forward and backward kernels (csrc/edgeconv.cu); the neighbour
    indices are an input (they carry no gradient, layers.py:33-35)."""

    @staticmethod
    def forward(ctx, x, idx32, k, w0, b0, w1, b1, w2, b2):
        B, C, N = x.shape
        out = torch.empty(B, 60, N, dtype=torch.float32, device=x.device)
        edgeconv_into(x, idx32, 0, k, [w0, w1, w2], [b0, b1, b2], out, ffma=True)
        ctx.k = k
        ctx.save_for_backward(x, idx32, w0, b0, w1, b1, w2, b2)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, idx32, w0, b0, w1, b1, w2, b2 = ctx.saved_tensors
        B, C, N = x.shape
        dy = dy.contiguous()
        dx = torch.zeros_like(x)
        w = [t.reshape(t.shape[0], t.shape[1]).contiguous() for t in (w0, w1, w2)]
        dw = [torch.zeros_like(t) for t in w]
        db = [torch.zeros_like(t) for t in (b0, b1, b2)]
        _lib.launch("pu3_edgeconv_bwd_f32", x, B, N, ctx.k, x.data_ptr(), C * N, idx32.data_ptr(), idx32.shape[2], 0,
                    w[0].data_ptr(), b0.data_ptr(), w[1].data_ptr(), b1.data_ptr(), w[2].data_ptr(), b2.data_ptr(),
                    dy.data_ptr(), 60 * N, dx.data_ptr(), C * N, dw[0].data_ptr(), db[0].data_ptr(), dw[1].data_ptr(),
                    db[1].data_ptr(), dw[2].data_ptr(), db[2].data_ptr())
        return (dx, None, None, dw[0].view_as(w0), db[0], dw[1].view_as(w1), db[1], dw[2].view_as(w2), db[2])


def dense_edge_conv(x, weights, biases, k, idx=None, max_group=None):
    """DenseEdgeConv.forward (layers.py:44-64).  x (B,C,N) -> (y (B,C+n*growth,N), idx (B,N,k) int64)."""
    _check_f32_cuda(x, "DenseEdgeConv")
    n_layers = len(weights)
    if _needs_grad(x, *weights, *biases) and _edgeconv_supported(x, weights) and k <= 32 and x.shape[2] <= 1500 \
            and native_edgeconv_backward:
        # training: fused forward + hand-written backward
        xc = x.contiguous()
        if idx is None:
            with torch.no_grad():
                _, idx_all, _ = _knn_raw(k + 1, xc.detach(), xc.detach(), True, max_group, want_knn=False, want_dist=False,
                                         idx_dtype=torch.int32)
            idx32 = idx_all[:, :, 1:].contiguous()
        else:
            idx32 = idx.to(torch.int32).contiguous()
        y = DenseEdgeConvFunction.apply(xc, idx32, k, weights[0], biases[0], weights[1], biases[1], weights[2], biases[2])
        return y, idx32.long()
    if _needs_grad(x, *weights, *biases) or not _edgeconv_supported(x, weights):
        # differentiable composition (same operators as the reference, group_knn on our kernels)
        if idx is None:
            knn_point, idx, _ = group_knn(k + 1, x, x, unique=True, max_group=max_group)
            idx = idx[:, :, 1:]
            knn_point = knn_point[:, :, :, 1:]
        else:
            B, C, N = x.shape
            knn_point = torch.gather(x.unsqueeze(2).expand(-1, -1, N, -1), 3, idx.unsqueeze(1).expand(-1, C, -1, -1))
        center = x.unsqueeze(-1).expand_as(knn_point)
        y = torch.cat([center, knn_point - center], dim=1)
        for i in range(n_layers):
            h = conv1x1_autograd(y, weights[i], biases[i])
            if i == 0:
                y = torch.cat([F.relu(h), x.unsqueeze(-1).expand(-1, -1, -1, k)], dim=1)
            elif i == n_layers - 1:
                y = torch.cat([h, y], dim=1)
            else:
                y = torch.cat([F.relu(h), y], dim=1)
        return y.max(dim=-1)[0], idx
    xc = x.contiguous()
    B, C, N = xc.shape
    if idx is None:
        _, idx32, _ = _knn_raw(k + 1, xc, xc, True, max_group, want_knn=False, want_dist=False, idx_dtype=torch.int32)
        off = 1
    else:
        idx32, off = idx.to(torch.int32).contiguous(), 0
    out = torch.empty(B, 60, N, dtype=torch.float32, device=x.device)
    edgeconv_into(xc, idx32, off, k, weights, biases, out)
    return out, idx32[:, :, off:off + k].long()


# ---- tensor-core (tcgen05, 3xTF32) 1x1 convolutions of the expansion head: csrc/conv_tc.cu -------------------------
def tc_supported(x, cout):
    """The TMA/tcgen05 path needs 16-byte aligned channel rows and cout <= 128."""
    return x.shape[2] % 4 == 0 and x.data_ptr() % 16 == 0 and cout <= 128 and (x.shape[0] == 1 or x.stride(0) % 4 == 0)


def tc_prepare(w, cin=None):
    """Split W (cout, >=cin) into the tf32 hi/lo shared-memory image the MMA reads (one small kernel)."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    if not w2.is_contiguous():
        w2 = w2.contiguous()
    cout, cin = w2.shape[0], (cin or w2.shape[1])
    nbytes = _lib.lib().pu3_conv_tc_wsplit_bytes(cin, cout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    _lib.launch("pu3_conv_tc_prepare_f32", w2, cin, cout, w2.data_ptr(), w2.shape[1], ws.data_ptr())
    return ws


def tc_conv_into(x, w, b, out, relu=False, wsplit=None):
    """out = act(W x + b) on the tensor cores; x, out: (B,C,N) channel slices as in conv_into."""
    B, Cin, N = x.shape
    Cout = out.shape[1]
    assert x.stride(2) == 1 and x.stride(1) == N and out.stride(2) == 1 and out.stride(1) == N
    ws = wsplit if wsplit is not None else tc_prepare(w)
    _lib.launch("pu3_conv_tc_f32", x, B, N, Cin, Cout, x.data_ptr(), x.stride(0) if B > 1 else Cin * N, ws.data_ptr(),
                _lib.ptr(b), out.data_ptr(), out.stride(0) if B > 1 else Cout * N, int(relu))
    return out


def tc_expand(x, w_full, b, code, r, wsplit=None, out=None):
    """Feature-expansion layer (upsampler.py:349-366): x (B,Cin,N), w_full (Cout, Cin+1) whose last column multiplies
    the 1-D code -> (B,Cout,N*r) = relu(W[:, :Cin] x + b + w_code * code[j])."""
    B, Cin, N = x.shape
    w2 = w_full.reshape(w_full.shape[0], w_full.shape[1])
    if not w2.is_contiguous():
        w2 = w2.contiguous()
    Cout = w2.shape[0]
    assert w2.shape[1] == Cin + 1 and x.stride(2) == 1 and x.stride(1) == N
    ws = wsplit if wsplit is not None else tc_prepare(w2, cin=Cin)
    if out is None:
        out = torch.empty(B, Cout, N * r, dtype=torch.float32, device=x.device)
    _lib.launch("pu3_conv_tc_expand_f32", x, B, N, Cin, Cout, r, x.data_ptr(), x.stride(0) if B > 1 else Cin * N,
                ws.data_ptr(), w2.data_ptr(), Cin + 1, Cin, _lib.ptr(b), code.data_ptr(), out.data_ptr(), Cout * N * r)
    return out


def tc_project(x, w_mid, b_mid, w_out, b_out, residual=None, res_div=1, wsplit=None):
    """Last two layers of the head (upsampler.py:369-372): relu(W_mid x + b_mid) stays on chip, then the <=3-channel
    projection (+ residual[..., p // res_div])."""
    B, Cin, N = x.shape
    wm = w_mid.reshape(w_mid.shape[0], w_mid.shape[1]).contiguous()
    wo = w_out.reshape(w_out.shape[0], w_out.shape[1]).contiguous()
    assert x.stride(2) == 1 and x.stride(1) == N
    ws = wsplit if wsplit is not None else tc_prepare(wm)
    out = torch.empty(B, wo.shape[0], N, dtype=torch.float32, device=x.device)
    rp, rbs, rn = None, 0, 1
    if residual is not None:
        assert residual.is_contiguous()
        rp, rbs, rn = residual.data_ptr(), residual.stride(0), residual.shape[2]
    _lib.launch("pu3_conv_tc_project_f32", x, B, N, Cin, wm.shape[0], wo.shape[0], x.data_ptr(),
                x.stride(0) if B > 1 else Cin * N, ws.data_ptr(), _lib.ptr(b_mid), wo.data_ptr(), _lib.ptr(b_out),
                out.data_ptr(), wo.shape[0] * N, rp, rbs, rn, res_div)
    return out


def tc_head(x, w1_full, b1, code, w2, b2, w3, b3, w4, b4, residual=None, wsplits=None):
    """The whole expansion head (upsampler.py:349-372) as one persistent tcgen05 kernel (csrc/head_tc.cu): x (B,Cin,N) features,
    w1_full (128, Cin+1) = up_layer1 incl. its code column, w2 (128,128), w3 (64,128), w4 (3,64), residual (B,3,N) -> (B,3,2N).
    Step ratio 2 only (code has two entries)."""
    B, Cin, N = x.shape
    m = lambda w: w.reshape(w.shape[0], w.shape[1]).contiguous()
    w1, w2, w3, w4 = m(w1_full), m(w2), m(w3), m(w4)
    assert w1.shape == (128, Cin + 1) and w2.shape == (128, 128) and w3.shape == (64, 128) and w4.shape == (3, 64)
    assert code.numel() == 2 and x.stride(2) == 1 and x.stride(1) == N
    ws = wsplits if wsplits is not None else (tc_prepare(w1, cin=Cin), tc_prepare(w2), tc_prepare(w3))
    out = torch.empty(B, 3, 2 * N, dtype=torch.float32, device=x.device)
    rp, rbs = None, 0
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == (B, 3, N)
        rp, rbs = residual.data_ptr(), 3 * N
    _lib.launch("pu3_head_tc_f32", x, B, N, Cin, x.data_ptr(), x.stride(0) if B > 1 else Cin * N, ws[0].data_ptr(),
                ws[1].data_ptr(), ws[2].data_ptr(), w1.data_ptr(), Cin + 1, Cin, _lib.ptr(b1), code.data_ptr(), _lib.ptr(b2),
                _lib.ptr(b3), w4.data_ptr(), _lib.ptr(b4), rp, rbs, out.data_ptr(), 3 * 2 * N)
    return out
