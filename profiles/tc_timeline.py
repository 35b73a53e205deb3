"""In-kernel timeline of conv_tc_kernel (CTA 0): SM cycle counter at the pipeline events of the first stages / tiles.
kinds: 0 producer passed empty-wait, 1 MMA warp saw TMA data, 2 converter done, 3 MMA warp saw converted data,
4 MMA warp issued commit, 5 MMA warp passed accumulator-free wait (tile), 6 epilogue saw accumulators (tile), 7 epilogue done."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
F = pu3.fused
dev = torch.device("cuda:0")
dbg = torch.zeros(8192, device=dev)
T = 1275
for (name, n, cin, cout) in [("up2 128->128", 624, 128, 128), ("fc1-like 128->64", 624, 128, 64)]:
    x = torch.rand(T, cin, n, device=dev); w = torch.rand(cout, cin, device=dev); ws = F.tc_prepare(w)
    out = torch.empty(T, cout, n, device=dev)
    for _ in range(2): F.tc_conv_into(x, w, None, out, wsplit=ws)
    torch.cuda.synchronize()
    pu3._lib.lib().pu3_conv_tc_set_debug(dbg.data_ptr()); dbg.zero_()
    F.tc_conv_into(x, w, None, out, wsplit=ws)
    torch.cuda.synchronize()
    pu3._lib.lib().pu3_conv_tc_set_debug(None)
    ev = dbg[4096:4096 + 8 * 256].view(torch.int32).cpu().view(8, 256).to(torch.int64) & 0xffffffff
    t0 = int(ev[0, 0])
    rel = lambda v: ((int(v) - t0) & 0xffffffff)
    print(f"== {name}: cycles relative to the first producer issue (1 us = ~1900 cycles)")
    print(" stage: produce  tma-seen  conv-done  mma-sees  commit")
    for s in range(0, 24):
        print(f"  {s:3d}: {rel(ev[0, s]):8d} {rel(ev[1, s]):8d} {rel(ev[2, s]):8d} {rel(ev[3, s]):8d} {rel(ev[4, s]):8d}")
    print(" tile: acc-free  epi-start  epi-done")
    for t in range(0, 6):
        print(f"  {t:3d}: {rel(ev[5, t]):8d} {rel(ev[6, t]):8d} {rel(ev[7, t]):8d}")
