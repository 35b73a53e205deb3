"""The "kernel to beat" table: the reference's own CUDA kernels
(sampling/sampling_cuda.cu:103-174 FPS, :26-41 gather, losses/nmdistance_cuda.cu:11-133,154-173 Chamfer) compiled for
sm_100a into oracle/_ref by oracle/build_ref.py, timed on the same B200 next to the new kernels at the live shapes of
the path.  MEASUREMENT INFRASTRUCTURE: runs in its own process (bench.py spawns it on rank 0 at N=1 and embeds the
JSON it prints as `kernels_to_beat`), so the product process never maps the reference's .so files.

    python profiles/kernels_to_beat.py            -> one JSON object on stdout
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from oracle import build_ref
    pu3 = importlib.import_module("3pu_pytorch_b200")
    rs, rl = build_ref.load()
    if rs is None:
        print(json.dumps({"unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    rows = []

    def row(op, shape, t_ref, t_new, alg_bytes, same):
        rows.append({"op": op, "shape": shape, "reference_ms": round(t_ref, 4), "ours_ms": round(t_new, 4),
                     "speedup": round(t_ref / t_new, 2), "ours_alg_gbs": round(alg_bytes / t_new / 1e6, 2),
                     "bit_identical": bool(same)})

    # ---- FPS at the shapes of the eval path (upsampler.py:78,158; main.py:228,379) ------------------------------
    for b, n, m, reps in [(1, 6240, 1248, 5), (1, 24960, 4992, 3), (32, 24960, 4992, 3), (32, 624, 10, 5), (1, 5000, 48, 5),
                          (1, 239616, 80000, 1)]:
        x = torch.rand(b, n, 3, generator=g).to(dev)
        want = torch.empty(b, m, dtype=torch.int32, device=dev)
        got = torch.empty(b, m, dtype=torch.int32, device=dev)
        t_r = torch.empty(b, n, device=dev)
        t_n = torch.empty(b, n, device=dev)

        def ref():
            t_r.fill_(1e10); rs.furthest_sampling(b, n, m, x, t_r, want)

        def new():
            t_n.fill_(1e10); pu3.sampling.furthest_sampling(b, n, m, x, t_n, got)
        w = 1 if n > 100000 else 2
        tr, tn = _time(ref, w, reps), _time(new, w, reps)
        row("furthest_sampling", f"b={b} n={n} m={m}", tr, tn, b * (12 * n + 4 * m + 8 * n), torch.equal(want, got))

    # ---- gather (operations.py:320 after every FPS; upsampler.py:58 seeds) -------------------------------------
    for b, c, n, m in [(32, 3, 24960, 4992), (32, 3, 624, 1), (32, 264, 312, 312)]:
        f = torch.rand(b, c, n, generator=g).to(dev)
        idx = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32).to(dev)
        o_r = torch.empty(b, c, m, device=dev); o_n = torch.empty(b, c, m, device=dev)
        tr = _time(lambda: rs.gather_forward(b, c, n, m, f, idx, o_r))
        tn = _time(lambda: pu3.sampling.gather_forward(b, c, n, m, f, idx, o_n))
        row("gather_forward", f"b={b} c={c} n={n} m={m}", tr, tn, b * (2 * c * m * 4 + 4 * m), torch.equal(o_r, o_n))
        go = torch.rand(b, c, m, generator=g).to(dev)
        g_r = torch.zeros(b, c, n, device=dev); g_n = torch.zeros(b, c, n, device=dev)
        tr = _time(lambda: rs.gather_backward(b, c, n, m, go, idx, g_r))
        tn = _time(lambda: pu3.sampling.gather_backward(b, c, n, m, go, idx, g_n))
        row("gather_backward", f"b={b} c={c} n={n} m={m}", tr, tn, b * (2 * c * m * 4 + 4 * m),
            torch.allclose(g_r, g_n, rtol=1e-5, atol=1e-5))

    # ---- Chamfer / NmDistance (model_loss.py:15,27) ------------------------------------------------------------------
    for b, n, m in [(32, 624, 624), (32, 4992, 4992), (256, 624, 624)]:
        x1 = torch.rand(b, n, 3, generator=g).to(dev); x2 = torch.rand(b, m, 3, generator=g).to(dev)
        outs = {}
        for name, mod in (("ref", rl), ("new", pu3.losses)):
            d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, dtype=torch.int32, device=dev)
            d2 = torch.empty(b, m, device=dev); i2 = torch.empty(b, m, dtype=torch.int32, device=dev)
            outs[name] = (d1, i1, d2, i2, _time(lambda: mod.nmdistance_forward(x1, x2, d1, d2, i1, i2)))
        same = all(torch.equal(a, c) for a, c in zip(outs["ref"][:4], outs["new"][:4]))
        row("nmdistance_forward", f"b={b} n={n} m={m}", outs["ref"][4], outs["new"][4], b * (n + m) * 20, same)
        g1 = torch.rand(b, n, generator=g).to(dev); g2 = torch.rand(b, m, generator=g).to(dev)
        res = {}
        for name, mod in (("ref", rl), ("new", pu3.losses)):
            gx1 = torch.zeros_like(x1); gx2 = torch.zeros_like(x2)
            _, i1, _, i2, _ = outs[name]

            def bwd():
                gx1.zero_(); gx2.zero_(); mod.nmdistance_backward(x1, x2, gx1, gx2, g1, g2, i1, i2)
            res[name] = (gx1, gx2, _time(bwd))
        same = torch.allclose(res["ref"][0], res["new"][0], rtol=1e-5, atol=1e-6) and \
            torch.allclose(res["ref"][1], res["new"][1], rtol=1e-5, atol=1e-6)
        row("nmdistance_backward", f"b={b} n={n} m={m}", res["ref"][2], res["new"][2], b * (n + m) * 32, same)

    print(json.dumps({"device": torch.cuda.get_device_name(0), "reference": "oracle/_ref (reference .cu, nvcc -O2, sm_100a)",
                      "timing": "CUDA events, mean of 3-5 launches after warm-up, outputs caller-allocated as in the reference ABI",
                      "rows": rows}))


if __name__ == "__main__":
    main()
