"""Bring-up diagnostics for csrc/conv_tc.cu: identity weights and index-coded activations show which
shared-memory element every (point, channel) of the MMA actually read."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
F = pu3.fused
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=220, sci_mode=False)
dbg = torch.zeros(4096, device=dev)
pu3._lib.lib().pu3_conv_tc_set_debug(dbg.data_ptr())
for variant in (int(v) for v in (sys.argv[1:] or ["2", "6", "8"])):
    pu3._lib.lib().pu3_conv_tc_set_variant(variant)
    for (N, Cin, Cout) in [(128, 32, 32)]:
        c = torch.arange(Cin, dtype=torch.float32).view(1, Cin, 1)
        p = torch.arange(N, dtype=torch.float32).view(1, 1, N)
        x = (c * 256 + p).to(dev).contiguous()
        w = torch.eye(Cout, Cin, device=dev)
        out = torch.full((1, Cout, N), -1.0, device=dev)
        dbg.zero_()
        F.tc_conv_into(x, w, None, out)
        torch.cuda.synchronize()
        y = out[0].cpu()
        want = x[0, :Cout].cpu()
        d = dbg.cpu()
        print(f"--- variant {variant} identity N={N} Cin={Cin} Cout={Cout}: exact {(y == want).float().mean():.3f}")
        print("tmem_base", hex(d[2048:2049].view(torch.int32).item()), "smem0", hex(d[2049:2050].view(torch.int32).item()))
        print("A raw row0 (32 floats):", d[0:32].tolist())
        print("A raw row1:", d[32:64].tolist())
        print("A raw row9:", d[9 * 32:10 * 32].tolist())
        print("B hi row0:", d[1024:1056].tolist())
        print("B hi row1:", d[1056:1088].tolist())
        print("B hi row5:", d[1024 + 160:1024 + 192].tolist())
        print("y[co=0..3, p=0..11]:\n", y[:4, :12])
        print("y[co=0, p=28..40]:", y[0, 28:41].tolist())
        print("y[:, p=0]:", y[:, 0].tolist())
        print("y[:, p=33]:", y[:, 33].tolist())
