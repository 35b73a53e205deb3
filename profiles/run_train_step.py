"""Profiling driver: ONE train step (B=32, ratio 16) between cudaProfilerStart/Stop (see run_step.py)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_net
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
net.load_state_dict(ref_net.make_params(4, seed=1), strict=True)
model = pu3.Model(net.to(dev), "train", lr_init=5e-4, weight_full_ratio=1.0)
model.use_cuda_graph = os.environ.get("GRAPH", "0") == "1"   # eager launches by default: every kernel is its own ncu result
g = torch.Generator().manual_seed(7)
x = torch.rand(32, 3, 312, generator=g).to(dev); gt = torch.rand(32, 3, 4992, generator=g).to(dev)
for _ in range(3):
    model.set_input(x, 16, label_pc=gt); model.optimize()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
torch.cuda.cudart().cudaProfilerStart()
model.set_input(x, 16, label_pc=gt); model.optimize()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("train step done; host issue %.1f ms, total %.1f ms" % (t_issue * 1e3, (time.perf_counter() - t0) * 1e3))
