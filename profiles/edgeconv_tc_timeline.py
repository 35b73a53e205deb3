"""(needs a library built with `make -C 3pu_pytorch_b200/csrc -B EXTRA=-DPU3_ET_TIMELINE`)
In-kernel timeline of the tensor-core edge-conv kernel: SM cycle counter of CTA (0,0), warpgroup 0, thread 0 at the phases of tiles 8..15."""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
pu3._lib.lib().pu3_edgeconv_set_tc(1)
lib = ctypes.CDLL(pu3._lib.LIB_PATH)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
ws = [(torch.randn(12, c, generator=g) * 0.25).to(dev) for c in (48, 36, 48)]
bs = [(torch.randn(12, generator=g) * 0.1).to(dev) for _ in range(3)]
b, n, k = 1275, 312, 32
x = torch.randn(b, 24, n, generator=g).to(dev)
idx = torch.randint(0, n, (b, n, k), generator=g).to(dev)
buf = torch.zeros(160, dtype=torch.int32, device=dev)
with torch.no_grad():
    pu3.fused.dense_edge_conv(x, ws, bs, k, idx=idx); torch.cuda.synchronize()
    lib.pu3_edgeconv_tc_set_timeline(ctypes.c_void_p(buf.data_ptr()))
    pu3.fused.dense_edge_conv(x, ws, bs, k, idx=idx); torch.cuda.synchronize()
    lib.pu3_edgeconv_tc_set_timeline(ctypes.c_void_p(0))
c = buf.cpu().to(torch.int64) & 0xffffffff
cn = ["entry", "weight images built", "prolog done (P | A of the cloud through the tensor core)", "warpgroup 0 finished its tiles", "CTA finished"]
print("CTA (0,0), cycles after kernel entry: " + "; ".join(f"{cn[i]} {int((c[i] - c[0]) & 0xffffffff)}" for i in range(1, 5)))
base = int(c[8])
rel = lambda v: int((int(v) - base) & 0xffffffff)
print("warpgroup 0, thread 0 and its MMA lane; cycles after layer 0 of tile 8 started")
print("tile: L0 start, L0 end | L1 start, wait passed, end | L2 start, wait passed, end || MMA lane: stage-1 ready seen, committed | stage-2 ready seen, committed")
for q in range(8):
    t = [rel(c[8 + q * 8 + i]) for i in range(8)]
    m = [rel(c[80 + q * 4 + i]) for i in range(4)]
    print(f"  {8 + q}: {t[0]}, {t[1]} | {t[2]}, {t[3]}, {t[4]} | {t[5]}, {t[6]}, {t[7]} || {m[0]}, {m[1]} | {m[2]}, {m[3]}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record(); pu3.fused.dense_edge_conv(x, ws, bs, k, idx=idx); e1.record(); torch.cuda.synchronize()
print(f"kernel (warm caches): {e0.elapsed_time(e1):.4f} ms")
