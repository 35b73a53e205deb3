"""DenseEdgeConv forward on the tensor cores (csrc/edgeconv_ts.cu and its first version csrc/edgeconv_tc.cu, tcgen05 3xTF32) -- network/layers.py:22-64 of the reference.
Checked against a float64 restatement of the reference's graph (edge feature [c, n - c], three 1x1 layers with dense
concatenation, max over k) at 1e-5, against the FFMA kernel, for bit-reproducibility, and on ragged shapes (points per cloud that
do not fill the 16-point blocks / 128-point prolog tiles, few and many clouds)."""
import ctypes
import pytest
import torch

pytestmark = pytest.mark.gpu
K = 32


def _reference64(x, idx, ws, bs):
    xd = x.double(); B, C, N = xd.shape
    nb = torch.gather(xd.unsqueeze(2).expand(B, C, N, N), 3, idx.unsqueeze(1).expand(B, C, N, K))
    c = xd.unsqueeze(3).expand(B, C, N, K)
    W = [w.double() for w in ws]; Bs = [b.double().view(1, -1, 1, 1) for b in bs]
    h0 = torch.relu(torch.einsum("oc,bcnk->bonk", W[0], torch.cat([c, nb - c], 1)) + Bs[0])          # layers.py:44-50
    h1 = torch.relu(torch.einsum("oc,bcnk->bonk", W[1], torch.cat([h0, c], 1)) + Bs[1])              # :51-57
    h2 = torch.einsum("oc,bcnk->bonk", W[2], torch.cat([h1, h0, c], 1)) + Bs[2]                      # :58-59 (no activation)
    return torch.cat([h2, h1, h0, c], 1).max(3)[0]                                                   # :62


def _weights(g, cuda, scale=0.25):
    ws = [(torch.randn(12, c, generator=g) * scale).to(cuda) for c in (48, 36, 48)]
    bs = [(torch.randn(12, generator=g) * 0.1).to(cuda) for _ in range(3)]
    return ws, bs


@pytest.mark.parametrize("b,n", [(1, 32), (2, 45), (5, 100), (3, 129), (3, 312), (40, 312), (7, 330), (300, 312), (2, 257)])
def test_tensor_core_edgeconv_against_float64_and_ffma(pu3, cuda, b, n):
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    g = torch.Generator().manual_seed(1000 * b + n)
    ws, bs = _weights(g, cuda)
    x = torch.randn(b, 24, n, generator=g).to(cuda)
    idx = torch.randint(0, n, (b, n, K), generator=g).to(cuda)
    outs = []
    try:
        for tc in (2, 1, 0):          # operands in tensor memory (shipped) / layer-1 operand images in shared memory / FFMA
            lib.pu3_edgeconv_set_tc(tc)
            with torch.no_grad():
                outs.append(pu3.fused.dense_edge_conv(x, ws, bs, K, idx=idx)[0].clone())
        lib.pu3_edgeconv_set_tc(2)
        with torch.no_grad():
            again = pu3.fused.dense_edge_conv(x, ws, bs, K, idx=idx)[0]
    finally:
        lib.pu3_edgeconv_set_tc(2)
    assert torch.equal(again, outs[0]), "the tensor-core kernel is not bit-reproducible"
    names = ("tensor-core (TS)", "tensor-core (SS + TS)", "FFMA")
    for name, o in zip(names, outs):
        assert torch.equal(o[:, 36:], x), f"{name}: pass-through channels must be copies"
    if b * n <= 4096:
        ref = _reference64(x, idx, ws, bs)
        for name, o in zip(names, outs):
            err = (o.double() - ref).abs()
            assert bool((err <= 1e-5 + 1e-5 * ref.abs()).all()), f"{name} kernel vs float64: max err {float(err.max()):.2e}"
    for name, o in zip(names[:2], outs[:2]):
        err = (o - outs[2]).abs()
        assert bool((err <= 1e-5 + 1e-5 * outs[2].abs()).all()), f"{name} vs FFMA: max diff {float(err.max()):.2e}"


def test_tensor_core_edgeconv_through_the_level_layout(pu3, cuda):
    """The call the level engine makes: neighbour lists of k + 1 with rank 0 dropped (idx_off = 1, layers.py:34-35), output written
    into a 60-channel slice of the 264-channel feature tensor (batch stride 264 n)."""
    g = torch.Generator().manual_seed(77)
    ws, bs = _weights(g, cuda)
    b, n = 6, 312
    x = torch.randn(b, 24, n, generator=g).to(cuda)
    idx33 = torch.randint(0, n, (b, n, K + 1), generator=g, dtype=torch.int32).to(cuda)
    feat = torch.full((b, 264, n), 7.0, device=cuda)
    pu3.fused.edgeconv_into(x, idx33, 1, K, ws, bs, feat[:, 100:160])
    ref = _reference64(x, idx33[:, :, 1:].long(), ws, bs)
    err = (feat[:, 100:160].double() - ref).abs()
    assert bool((err <= 1e-5 + 1e-5 * ref.abs()).all()), f"max err {float(err.max()):.2e}"
    assert bool((feat[:, :100] == 7.0).all()) and bool((feat[:, 160:] == 7.0).all()), "wrote outside its 60 channels"


def test_train_mode_forward_stays_on_the_ffma_kernels(pu3, cuda):
    """A forward that will be differentiated returns exactly what pu3_edgeconv_bwd_f32 recomputes (pu3_edgeconv_ffma_f32)."""
    lib = ctypes.CDLL(pu3._lib.LIB_PATH)
    g = torch.Generator().manual_seed(5)
    ws, bs = _weights(g, cuda)
    x = torch.randn(2, 24, 312, generator=g).to(cuda)
    idx = torch.randint(0, 312, (2, 312, K), generator=g).to(cuda)
    xg = x.clone().requires_grad_()
    y_train = pu3.fused.dense_edge_conv(xg, ws, bs, K, idx=idx)[0]
    lib.pu3_edgeconv_set_tc(0)
    try:
        with torch.no_grad():
            y_ffma = pu3.fused.dense_edge_conv(x, ws, bs, K, idx=idx)[0]
    finally:
        lib.pu3_edgeconv_set_tc(2)
    assert y_train.requires_grad and torch.equal(y_train.detach(), y_ffma)
