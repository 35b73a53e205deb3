"""Bring-up probe of the tcgen05 3xTF32 convolution kernels (csrc/conv_tc.cu): accuracy against a float64
reference for every epilogue mode and a few ragged shapes, then timing at the level-4 shapes of the B=32 step.
Each descriptor variant runs in its own process under a timeout (a protocol error traps, it cannot hang).
    python profiles/tc_probe.py            # all variants
    python profiles/tc_probe.py --variant 0
"""
import argparse
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(got, want):
    return float((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-30))


def worst_ratio(got, want, rtol=1e-5, atol=1e-5):
    """max |got - want| / (atol + rtol |want|): the element-wise bar of tests/ (must stay below 1)"""
    return float(((got.double() - want).abs() / (atol + rtol * want.abs())).max())


def run(variant, timing, sweep=False):
    import torch
    pu3 = importlib.import_module("3pu_pytorch_b200")
    F = pu3.fused
    pu3._lib.lib().pu3_conv_tc_set_variant(variant)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    ok = True

    def rnd(*s):
        return (torch.rand(*s, generator=g) * 2 - 1).to(dev)

    # plain
    for (B, N, Cin, Cout, relu) in [(3, 624, 128, 128, True), (2, 312, 264, 128, False), (8, 312, 264, 128, False), (1, 40, 84, 24, True),
                                    (5, 100, 8, 64, False), (3, 128, 32, 128, False), (301, 624, 128, 128, True)]:
        x, w, b = rnd(B, Cin, N), rnd(Cout, Cin) * 0.2, rnd(Cout)
        out = torch.full((B, Cout, N), float("nan"), device=dev)
        F.tc_conv_into(x, w, b, out, relu=relu)
        torch.cuda.synchronize()
        ref = torch.matmul(w.double(), x.double()) + b.double().view(1, -1, 1)
        ref = ref.clamp_min(0) if relu else ref
        e = rel_err(out, ref)
        ffma = torch.empty_like(out); F.conv_into(x, w, b, ffma, relu=relu)
        e2 = rel_err(ffma, ref)
        print(f"variant {variant} plain   B={B} N={N} {Cin}->{Cout} relu={relu}: rel err {e:.3e} (FFMA kernel {e2:.3e}), element-wise worst ratio "
              f"{worst_ratio(out, ref):.2f} (FFMA {worst_ratio(ffma, ref):.2f}) nan={bool(torch.isnan(out).any())}")
        ok &= e < 2e-6
    # slices of a bigger buffer (the way the level engine calls it)
    B, N = 4, 312
    feat = rnd(B, 264, N)
    w, b = rnd(24, 204) * 0.2, rnd(24)
    out = torch.empty(B, 24, N, device=dev)
    F.tc_conv_into(feat[:, 60:], w, b, out, relu=True)
    ref = (torch.matmul(w.double(), feat[:, 60:].double()) + b.double().view(1, -1, 1)).clamp_min(0)
    e = rel_err(out, ref); print(f"variant {variant} slice   204->24: rel err {e:.3e}"); ok &= e < 2e-6
    # expand
    for (B, N, Cin, Cout, r) in [(3, 312, 264, 128, 2), (2, 100, 40, 64, 3)]:
        x, w, b = rnd(B, Cin, N), rnd(Cout, Cin + 1) * 0.2, rnd(Cout)
        code = torch.linspace(-0.2, 0.2, r, device=dev)
        out = F.tc_expand(x, w, b, code, r)
        torch.cuda.synchronize()
        pre = torch.matmul(w[:, :Cin].double(), x.double()) + b.double().view(1, -1, 1)
        ref = (pre.unsqueeze(-1) + (w[:, Cin].double().view(1, -1, 1, 1) * code.double().view(1, 1, 1, -1))).clamp_min(0).reshape(B, Cout, N * r)
        e = rel_err(out, ref); print(f"variant {variant} expand  B={B} N={N} {Cin}->{Cout} r={r}: rel err {e:.3e}"); ok &= e < 2e-6
    # project
    for (B, N, Cin, Cmid, Cout, div) in [(3, 624, 128, 64, 3, 2), (2, 52, 16, 40, 2, 1)]:
        x, wm, bm, wo, bo = rnd(B, Cin, N), rnd(Cmid, Cin) * 0.2, rnd(Cmid), rnd(Cout, Cmid) * 0.2, rnd(Cout)
        res = rnd(B, Cout, N // div)
        out = F.tc_project(x, wm, bm, wo, bo, residual=res, res_div=div)
        torch.cuda.synchronize()
        h = (torch.matmul(wm.double(), x.double()) + bm.double().view(1, -1, 1)).clamp_min(0)
        ref = torch.matmul(wo.double(), h) + bo.double().view(1, -1, 1) + res.double().repeat_interleave(div, dim=2)
        e = rel_err(out, ref); print(f"variant {variant} project B={B} N={N} {Cin}->{Cmid}->{Cout}: rel err {e:.3e}"); ok &= e < 2e-6
    print(f"variant {variant}: {'ALL OK' if ok else 'MISMATCH'}")
    if sweep:
        ok = True
    if ok and timing:
        T = 1275
        feat = rnd(T, 264, 312)
        w1, b1, code = rnd(128, 265) * 0.1, rnd(128), torch.tensor([-0.2, 0.2], device=dev)
        w2, b2 = rnd(128, 128) * 0.1, rnd(128)
        w3, b3, w4, b4 = rnd(64, 128) * 0.1, rnd(64), rnd(3, 64) * 0.1, rnd(3)
        xyz = rnd(T, 3, 312)
        s1, s2, s3 = F.tc_prepare(w1, cin=264), F.tc_prepare(w2), F.tc_prepare(w3)
        h2 = torch.empty(T, 128, 624, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn, reps=5):
            ts = []
            for _ in range(reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sorted(ts[2:])[len(ts[2:]) // 2]

        h1 = F.tc_expand(feat, w1, b1, code, 2, wsplit=s1)
        if sweep:   # which part of the pipeline bounds the kernel: 2 = one MMA instead of three, 16 = no stores, 32 = no MMA
            for v in (0, 64, 2, 16, 32, 48, 48 + 128, 48 + 256, 48 + 512, 48 + 128 + 256, 48 + 128 + 512):   # 64 = no L2 prefetch, 128 = no conversion, 256 = no weight copy, 512 = no activation TMA
                pu3._lib.lib().pu3_conv_tc_set_variant(v)
                t1 = timed(lambda: F.tc_expand(feat, w1, b1, code, 2, wsplit=s1))
                t2 = timed(lambda: F.tc_conv_into(h1, w2, b2, h2, relu=True, wsplit=s2))
                t3 = timed(lambda: F.tc_project(h2, w3, b3, w4, b4, residual=xyz, res_div=2, wsplit=s3))
                print(f"sweep variant {v:3d}: up1+expand {t1:.3f} ms   up2 {t2:.3f} ms   fc1+fc2 {t3:.3f} ms")
            pu3._lib.lib().pu3_conv_tc_set_variant(0)
            return True
        t1 = timed(lambda: F.tc_expand(feat, w1, b1, code, 2, wsplit=s1))
        t2 = timed(lambda: F.tc_conv_into(h1, w2, b2, h2, relu=True, wsplit=s2))
        t3 = timed(lambda: F.tc_project(h2, w3, b3, w4, b4, residual=xyz, res_div=2, wsplit=s3))
        # the FFMA kernels they replace
        pre = torch.empty(T, 128, 312, device=dev); h3 = torch.empty(T, 64, 624, device=dev); o = torch.empty(T, 3, 624, device=dev)
        f1 = timed(lambda: F.conv_into(feat, w1[:, :264].contiguous(), b1, pre))
        f2 = timed(lambda: F.conv_into(h1, w2, b2, h2, relu=True))
        f3 = timed(lambda: F.conv_into(h2, w3, b3, h3, relu=True))
        f4 = timed(lambda: F.conv_into(h3, w4, b4, o, residual=xyz, res_div=2))
        gf1, gf2, gf3 = 2 * 264 * 128 * T * 312 / 1e9, 2 * 128 * 128 * T * 624 / 1e9, 2 * 128 * 64 * T * 624 / 1e9
        print(f"timing T={T}: up1+expand {t1:.3f} ms ({gf1 / t1:.1f} fp32-equivalent TFLOP/s; FFMA conv alone {f1:.3f} ms)")
        print(f"timing T={T}: up2        {t2:.3f} ms ({gf2 / t2:.1f} TFLOP/s; FFMA {f2:.3f} ms)")
        print(f"timing T={T}: fc1+fc2    {t3:.3f} ms ({gf3 / t3:.1f} TFLOP/s; FFMA {f3:.3f} + {f4:.3f} ms)")
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--no-timing", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    a = ap.parse_args()
    if a.variant is not None:
        sys.exit(0 if run(a.variant, not a.no_timing, a.sweep) else 1)
    for v in (0, 1):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", str(v)], timeout=240)
            print(f"== variant {v}: exit {r.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"== variant {v}: TIMEOUT", flush=True)
