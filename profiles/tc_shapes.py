"""Timing of the tcgen05 convolution over shapes that isolate the point tail, the channel tail and the cloud size."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
F = pu3.fused
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]


T = 1275
for (n, cin, cout) in [(312, 264, 128), (312, 256, 128), (384, 264, 128), (624, 264, 128), (312, 128, 128), (256, 256, 128),
                       (624, 128, 128), (640, 128, 128), (624, 128, 64)]:
    x = torch.rand(T, cin, n, device=dev)
    w = torch.rand(cout, cin, device=dev)
    ws = F.tc_prepare(w)
    out = torch.empty(T, cout, n, device=dev)
    res = []
    for v in (0, 48):
        pu3._lib.lib().pu3_conv_tc_set_variant(v)
        res.append(timed(lambda: F.tc_conv_into(x, w, None, out, wsplit=ws)))
    pu3._lib.lib().pu3_conv_tc_set_variant(0)
    spc = (n + 127) // 128
    tiles = (T * spc + 1) // 2
    stages = -(-tiles // 148) * ((cin + 31) // 32)
    print(f"n={n} cin={cin} cout={cout}: full {res[0]:.3f} ms, load+convert only {res[1]:.3f} ms  ({res[1] * 1e3 / stages:.2f} us/stage, "
          f"in {T * cin * n * 4 / 1e6:.0f} MB -> {T * cin * n * 4 / res[1] / 1e9:.2f} TB/s)")

# expansion epilogue: pre-allocated output vs allocation inside the timed region
n, cin, cout = 312, 264, 128
x = torch.rand(T, cin, n, device=dev)
w = torch.rand(cout, cin + 1, device=dev)
b = torch.rand(cout, device=dev)
code = torch.tensor([-0.2, 0.2], device=dev)
ws = F.tc_prepare(w, cin=cin)
out = torch.empty(T, cout, n * 2, device=dev)
for v in (0, 16, 48):
    pu3._lib.lib().pu3_conv_tc_set_variant(v)
    t_pre = timed(lambda: F.tc_expand(x, w, b, code, 2, wsplit=ws, out=out))
    t_alloc = timed(lambda: F.tc_expand(x, w, b, code, 2, wsplit=ws))
    print(f"expand variant {v}: preallocated {t_pre:.3f} ms, allocating {t_alloc:.3f} ms")
pu3._lib.lib().pu3_conv_tc_set_variant(0)
