"""Profiling driver: warm up, then run ONE eval step (B=32, 312 -> 4992) between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/run_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernel> -c 3 \
        -o gpurun_out/prof python profiles/run_step.py
Numbers printed under ncu are never bench values."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_net  # noqa: E402

pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
net.load_state_dict(ref_net.make_params(4, seed=1), strict=True)
net = net.to(dev).eval()
net.use_cuda_graph = os.environ.get("GRAPH", "0") == "1"   # eager launches by default: every kernel is its own ncu result
x = bench.make_inputs(0).to(dev)
with torch.no_grad():
    for _ in range(int(os.environ.get("WARMUP", "2"))):
        net(x, ratio=16)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    y = net(x, ratio=16)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("step done", tuple(y.shape))
