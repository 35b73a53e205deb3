"""DenseEdgeConv forward: tensor-core kernel (csrc/edgeconv_tc.cu) against the FFMA kernel (csrc/edgeconv.cu) at the level
shapes of the B=32 eval step -- largest difference and time per launch.  Usage: python profiles/edgeconv_tc_check.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
F = pu3.fused
lib = pu3._lib.lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]


g = torch.Generator().manual_seed(3)
ws = [(torch.randn(12, c, generator=g) * 0.25).to(dev) for c in (48, 36, 48)]
bs = [(torch.randn(12, generator=g) * 0.1).to(dev) for _ in range(3)]
k = 32
tot = [0.0, 0.0]


def reference64(x, idx):
    """network/layers.py:22-64 in float64: edge feature [c, n - c], three 1x1 layers with dense concatenation, max over k."""
    xd = x.double(); B, C, N = xd.shape
    nb = torch.gather(xd.unsqueeze(2).expand(B, C, N, N), 3, idx.unsqueeze(1).expand(B, C, N, k))      # (B,C,N,k): x[:, :, idx]
    c = xd.unsqueeze(3).expand(B, C, N, k)
    W = [w.double() for w in ws]; Bs = [b.double().view(1, -1, 1, 1) for b in bs]
    h0 = torch.relu(torch.einsum("oc,bcnk->bonk", W[0], torch.cat([c, nb - c], 1)) + Bs[0])
    h1 = torch.relu(torch.einsum("oc,bcnk->bonk", W[1], torch.cat([h0, c], 1)) + Bs[1])
    h2 = torch.einsum("oc,bcnk->bonk", W[2], torch.cat([h1, h0, c], 1)) + Bs[2]
    return torch.cat([h2, h1, h0, c], 1).max(3)[0]
for b, n in [(2, 45), (5, 100), (3, 312), (32, 312), (160, 312), (640, 312), (1275, 312), (7, 330)]:
    x = torch.randn(b, 24, n, generator=g).to(dev)
    idx = torch.randint(0, n, (b, n, k), generator=g).to(dev)
    outs = []
    for tc in (2, 0):
        lib.pu3_edgeconv_set_tc(tc)
        with torch.no_grad():
            outs.append(F.dense_edge_conv(x, ws, bs, k, idx=idx)[0].clone())
    torch.cuda.synchronize()
    lib.pu3_edgeconv_set_tc(2)
    with torch.no_grad():
        same = all(torch.equal(F.dense_edge_conv(x, ws, bs, k, idx=idx)[0], outs[0]) for _ in range(6))
    if not same:
        print("    !!! the tensor-core kernel is not bit-reproducible on this shape")
    if b * n <= 32 * 312:
        ref = reference64(x, idx)
        e = [(o.double() - ref).abs() for o in outs]
        print(f"    against float64: tensor-core max {float(e[0].max()):.2e} mean {float(e[0].mean()):.2e};  FFMA max {float(e[1].max()):.2e} mean {float(e[1].mean()):.2e}"
              f"  (max |y| {float(ref.abs().max()):.1f})")
    d = (outs[0] - outs[1]).abs()
    tol = 1e-5 + 1e-5 * outs[1].abs()
    bad = int((d > tol).sum())
    per = [float(d[:, a:a + 12].max()) for a in (0, 12, 24)] + [float(d[:, 36:].max())]
    t = []
    for tc in (2, 0):
        lib.pu3_edgeconv_set_tc(tc)
        with torch.no_grad():
            t.append(timed(lambda: F.dense_edge_conv(x, ws, bs, k, idx=idx)))
    if n == 312 and b >= 32:
        tot[0] += t[0]; tot[1] += t[1]
    print(f"b={b:5d} n={n}: max|tc - ffma| h2/h1/h0/centre = {per[0]:.2e} {per[1]:.2e} {per[2]:.2e} {per[3]:.1e}, outside 1e-5: {bad} of {d.numel()};"
          f"  tensor-core {t[0]:.4f} ms, FFMA {t[1]:.4f} ms ({t[1] / t[0]:.2f}x)", flush=True)
lib.pu3_edgeconv_set_tc(2)
print(f"sum over the four level shapes: tensor-core {tot[0]:.3f} ms, FFMA {tot[1]:.3f} ms (x4 blocks per level in the step)")
