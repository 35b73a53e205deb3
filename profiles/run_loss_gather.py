"""Profiling driver: the Chamfer / NmDistance kernels at the train shape (32,624,3)^2 and at the eval-output shape, and the gather
kernels at their live shapes, between cudaProfilerStart/Stop (ncu --profile-from-start off ...).  Numbers printed under ncu are
never bench values."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def chamfer(b, n, m):
    x1 = torch.rand(b, n, 3, generator=g).to(dev); x2 = torch.rand(b, m, 3, generator=g).to(dev)
    d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, dtype=torch.int32, device=dev)
    d2 = torch.empty(b, m, device=dev); i2 = torch.empty(b, m, dtype=torch.int32, device=dev)
    pu3.losses.nmdistance_forward(x1, x2, d1, d2, i1, i2)
    gx1 = torch.zeros_like(x1); gx2 = torch.zeros_like(x2)
    g1 = torch.rand(b, n, generator=g).to(dev); g2 = torch.rand(b, m, generator=g).to(dev)
    pu3.losses.nmdistance_backward(x1, x2, gx1, gx2, g1, g2, i1, i2)


def gather(b, c, n, m):
    f = torch.rand(b, c, n, generator=g).to(dev)
    idx = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32).to(dev)
    o = torch.empty(b, c, m, device=dev)
    pu3.sampling.gather_forward(b, c, n, m, f, idx, o)
    gp = torch.zeros(b, c, n, device=dev)
    pu3.sampling.gather_backward(b, c, n, m, o, idx, gp)


shapes_c = [(32, 624, 624), (256, 624, 624), (32, 4992, 4992)]
shapes_g = [(32, 3, 24960, 4992), (32, 264, 312, 312)]
for s in shapes_c: chamfer(*s)
for s in shapes_g: gather(*s)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for s in shapes_c: chamfer(*s)
for s in shapes_g: gather(*s)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
