"""How sensitive is the end-to-end 16x cloud comparison to rounding-level changes?  Prints cloud_match_fraction
(tol 1e-4) of Net.forward against the CPU oracle for a few inputs, with the fast and the generic edge-conv kernel
(they differ by a few ulp in layer 0) and with the tcgen05 / FFMA expansion head."""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_net
from tests.util import cloud_match_fraction
pu3 = importlib.import_module("3pu_pytorch_b200")
lib = ctypes.CDLL(pu3._lib.LIB_PATH)
dev = torch.device("cuda:0")
params = ref_net.make_params(4, seed=1)
net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
net.load_state_dict(params, strict=True)
net = net.to(dev).eval()
for seed in (13, 14, 15):
    g = torch.Generator().manual_seed(seed)
    x = ref_net.normalize_point_batch(torch.rand(1, 3, 312, generator=g))[0]
    with torch.no_grad():
        want = ref_net.net_forward(params, x, ratio=16, max_up_ratio=16)
        outs = {}
        for name, gen, tc in (("fast-ec + tc-head", 0, 1), ("generic-ec + tc-head", 1, 1), ("fast-ec + ffma-head", 0, 0), ("generic-ec + ffma-head", 1, 0)):
            lib.pu3_edgeconv_force_generic(gen); lib.pu3_level_set_tc(tc)
            outs[name] = net(x.to(dev), ratio=16).cpu()
        lib.pu3_edgeconv_force_generic(0); lib.pu3_level_set_tc(1)
    d = torch.cdist(outs["fast-ec + tc-head"][0].t().double(), want[0].t().double())
    print(f"seed {seed}: " + ", ".join(f"{k}: {cloud_match_fraction(v[0], want[0], tol=1e-4):.4f}" for k, v in outs.items()) +
          f" | chamfer(fast,oracle) mean NN dist {d.min(1)[0].mean():.2e} / {d.min(0)[0].mean():.2e}, oracle point spacing {torch.cdist(want[0].t().double(), want[0].t().double()).topk(2, largest=False)[0][:, 1].mean():.2e}")
