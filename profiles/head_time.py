"""Fused tcgen05 head (csrc/head_tc.cu) against the three-kernel tcgen05 head at the four level shapes of the B=32 eval step,
plus the in-kernel event timeline of CTA 0 (cycles per tile phase).  Usage: python profiles/head_time.py [--timeline] [--only=T]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
F = pu3.fused
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


g = torch.Generator().manual_seed(0)
cin, n = 264, 312
w1 = (torch.randn(128, cin + 1, generator=g) * 0.08).to(dev); b1 = torch.randn(128, generator=g).mul(0.1).to(dev)
w2 = (torch.randn(128, 128, generator=g) * 0.12).to(dev); b2 = torch.randn(128, generator=g).mul(0.1).to(dev)
w3 = (torch.randn(64, 128, generator=g) * 0.12).to(dev); b3 = torch.randn(64, generator=g).mul(0.1).to(dev)
w4 = (torch.randn(3, 64, generator=g) * 0.2).to(dev); b4 = torch.randn(3, generator=g).mul(0.1).to(dev)
code = torch.tensor([-1.0, 1.0], device=dev)
ws = (F.tc_prepare(w1, cin=cin), F.tc_prepare(w2), F.tc_prepare(w3))
total_f = total_3 = 0.0
mode = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--mode=")]
pu3._lib.lib().pu3_head_tc_set_mode(mode[0] if mode else 2)
print("mode:", {0: "operands in shared memory (SS)", 1: "operands in tensor memory (TS)", 2: "CTA pairs, operands in tensor memory (cta_group::2, TS)"}[mode[0] if mode else 2])
only = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--only=")]
for T in (only or (32, 160, 640, 1275)):
    x = torch.randn(T, cin, n, device=dev)
    res = torch.randn(T, 3, n, device=dev)
    h1 = torch.empty(T, 128, 2 * n, device=dev); h2 = torch.empty_like(h1)

    def three():
        F.tc_expand(x, w1, b1, code, 2, wsplit=ws[0], out=h1)
        F.tc_conv_into(h1, w2, b2, h2, relu=True, wsplit=ws[1])
        return F.tc_project(h2, w3, b3, w4, b4, residual=res, res_div=2, wsplit=ws[2])

    fused = lambda: F.tc_head(x, w1, b1, code, w2, b2, w3, b3, w4, b4, residual=res, wsplits=ws)
    a, b = fused(), three()
    tf, t3 = timed(fused), timed(three)
    total_f += tf; total_3 += t3
    tiles = -(-T * 10 // 4)
    per_sm = -(-tiles // 148)
    mma_cycles = 33 * 192 + 16 * 2 * 192 + 16 * 2 * 96            # tensor-pipe cycles per tile at the nominal tf32 rate
    print(f"T={T:5d}: fused {tf:.4f} ms, three kernels {t3:.4f} ms ({t3 / tf:.2f}x); {tiles} tiles, {per_sm} per SM, "
          f"{tf * 1e3 / per_sm:.2f} us per tile (tensor work {mma_cycles} cycles); max |fused - three| {float((a - b).abs().max()):.2e}")
print(f"sum over the four levels: fused {total_f:.3f} ms, three kernels {total_3:.3f} ms")

if "--timeline" in sys.argv:
    T = 1275
    x = torch.randn(T, cin, n, device=dev); res = torch.randn(T, 3, n, device=dev)
    buf = torch.zeros(24 * 512, dtype=torch.int32, device=dev)
    lib = pu3._lib.lib()
    lib.pu3_head_tc_set_debug(buf.data_ptr())
    F.tc_head(x, w1, b1, code, w2, b2, w3, b3, w4, b4, residual=res, wsplits=ws)
    torch.cuda.synchronize()
    lib.pu3_head_tc_set_debug(None)
    ev = buf.cpu().view(24, 512).numpy().astype("int64") & 0xffffffff
    m = mode[0] if mode else 2
    if m == 0:
        names = {0: "P1 start", 1: "P1 issued", 2: "P2 start", 3: "P2 issued", 4: "P3 issued", 5: "E1 start", 6: "E2 start", 7: "E3 start", 8: "E3 done"}
    elif m == 1:
        names = {0: "up1 start", 1: "up1 issued", 2: "up2(0) start", 3: "up2(0) issued", 4: "fc1(0) issued", 5: "up2(1) start",
                 6: "up2(1) issued", 7: "fc1(1) issued", 8: "E1(0) start", 9: "E2(0) start", 10: "E3(0) start", 11: "E1(1) start",
                 12: "E2(1) start", 13: "E3(1) start", 14: "E3(0) done", 15: "E3(1) done"}
    else:
        names = {0: "up1 start", 1: "up1 issued", 2: "up2(0) start", 3: "up2(0) issued", 4: "up2(1) start", 5: "up2(1) issued",
                 6: "fc1(0) issued", 7: "fc1(1) issued", 8: "E1(0) start", 9: "E2(0) start", 10: "E3(0) start", 11: "E1(1) start",
                 12: "E2(1) start", 13: "E3(1) start", 14: "E3(0) done", 15: "E3(1) done"}
    t0 = ev[0, 2]
    for it in range(2, 6):
        print(f"tile {it}: " + " | ".join(f"{names[k]} {int((ev[k, it] - t0) & 0xffffffff)}" for k in sorted(names)))
    per = [(ev[0, it + 1] - ev[0, it]) & 0xffffffff for it in range(2, 18)]
    print("cycles per tile (up1 start to up1 start):", [int(v) for v in per])
    if not (mode and mode[0] == 0):
        for k, nm in ((16, "MMA warp waits for weights"), (17, "MMA warp waits for staged up1 operands"), (18, "MMA warp waits for staged up2/fc1 operands"),
                      (19, "MMA warp waits for epilogue events"), (20, "converter waits for the raw TMA block"), (21, "converter waits for a staging slot")):
            print(f"{nm}: cycles per tile", [int(v) for v in ev[k, 2:10]])
        base = ev[0, 3]
        rel = lambda v: int((v - base) & 0xffffffff)
        if m == 1:
            print("tile 3, epilogue group 0 / quarter 0, cycles after up1 start; per chunk: start | TMEM loaded | math done | stores issued | arrived")
            for ph, nm in ((0, "E1(0)"), (20, "E2(0)")):
                for ch in range(4):
                    print(f"  {nm} chunk {ch}: " + " | ".join(str(rel(ev[22, ph + ch * 5 + k])) for k in range(5)))
        print("  MMA warp saw operands ready: up2(0)", [rel(ev[23, c]) for c in range(4)], "fc1(0)", [rel(ev[23, 4 + c]) for c in range(4)],
              "up2(1)", [rel(ev[23, 8 + c]) for c in range(4)], "fc1(1)", [rel(ev[23, 12 + c]) for c in range(4)])
