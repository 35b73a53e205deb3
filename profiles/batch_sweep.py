"""Step time vs number of requests in the batch (host-launch floor vs GPU work), eval 312->4992."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import ref_net
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
net.load_state_dict(ref_net.make_params(4, seed=1), strict=True)
net = net.to(dev).eval(); net.eval_groups = 1
for B in (1, 4, 8, 16, 32, 64):
    x = bench.make_inputs(0, B).to(dev)
    with torch.no_grad():
        for _ in range(3): net(x, ratio=16)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(5): net(x, ratio=16)
        e1.record(); t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
    print(f"B={B:3d}  gpu {e0.elapsed_time(e1)/5:8.2f} ms/step  host-issue {t_host/5*1e3:8.2f} ms/step  {B/(e0.elapsed_time(e1)/5e3):8.1f} patches/s")
