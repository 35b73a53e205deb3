import importlib, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, os.getcwd())
import torch, bench
from oracle import ref_net
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
net.load_state_dict(ref_net.make_params(4, seed=1), strict=True); net = net.to(dev).eval()
x = bench.make_inputs(0).to(dev)
for g in (False, True):
    net.use_cuda_graph = g
    with torch.no_grad():
        net(x, ratio=16); net(x, ratio=16); torch.cuda.synchronize()
        prof = pu3._lib.Profiler(timing=False); prev = pu3._lib.set_profiler(prof)
        net(x, ratio=16); torch.cuda.synchronize()
        pu3._lib.set_profiler(prev)
    print("graph" if g else "eager", prof.launches, dict(prof.calls))
