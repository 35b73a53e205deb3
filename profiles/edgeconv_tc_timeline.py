"""(needs a library built with `make -C 3pu_pytorch_b200/csrc -B EXTRA=-DPU3_ET_TIMELINE`)
In-kernel timeline of the tensor-core edge-conv kernel: SM cycle counter of CTA (0,0), warpgroup 0, thread 0 at the phases of tiles 8..15."""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
lib = ctypes.CDLL(pu3._lib.LIB_PATH)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
ws = [(torch.randn(12, c, generator=g) * 0.25).to(dev) for c in (48, 36, 48)]
bs = [(torch.randn(12, generator=g) * 0.1).to(dev) for _ in range(3)]
b, n, k = 1275, 312, 32
x = torch.randn(b, 24, n, generator=g).to(dev)
idx = torch.randint(0, n, (b, n, k), generator=g).to(dev)
buf = torch.zeros(8 * 16 + 32, dtype=torch.int32, device=dev)
names = ["start", "layer 0 done", "barrier passed", "MMAs issued", "stage-1 wait passed", "TMEM loaded", "layer-1 epilogue done", "barrier passed",
         "stage-2 wait passed", "TMEM loaded", "tile end"]
with torch.no_grad():
    pu3.fused.dense_edge_conv(x, ws, bs, k, idx=idx); torch.cuda.synchronize()
    lib.pu3_edgeconv_tc_set_timeline(ctypes.c_void_p(buf.data_ptr()))
    pu3.fused.dense_edge_conv(x, ws, bs, k, idx=idx); torch.cuda.synchronize()
    lib.pu3_edgeconv_tc_set_timeline(ctypes.c_void_p(0))
full = buf.cpu().to(torch.int64) & 0xffffffff
t = full[:128].view(8, 16)
print("cycles after the tile's start: " + " | ".join(names[1:]))
for r in range(7):
    row = [int((t[r, ph] - t[r, 0]) & 0xffffffff) for ph in range(1, 11)]
    print(f"  tile {8 + r}: {row}   period {int((t[r + 1, 0] - t[r, 0]) & 0xffffffff)}")
c = full[128:]
cn = ["entry", "barriers + TMEM allocated", "cloud + centre weights loaded (issued)", "weight images built", "barrier", "P_j done", "centre terms done", "centre copy done",
      "barrier", "operand images zeroed, main loop starts", "warpgroup 0 finished", "CTA finished"]
print("CTA (0,0), cycles after kernel entry: " + "; ".join(f"{cn[i]} {int((c[i] - c[0]) & 0xffffffff)}" for i in range(1, 12)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record(); pu3.fused.dense_edge_conv(x, ws, bs, k, idx=idx); e1.record(); torch.cuda.synchronize()
print(f"kernel (warm caches): {e0.elapsed_time(e1):.4f} ms; {b} CTAs on 296 slots")
