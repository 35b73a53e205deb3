"""Device time of furthest_point_sample (transpose + FPS + gather) on the merge-FPS shapes of the B=32 eval step."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
for (n, m) in [(6240, 1248), (12480, 2496), (24960, 4992)]:
    xyz = torch.rand(32, 3, n, device=dev)
    for _ in range(2): pu3.operations.furthest_point_sample(xyz, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): pu3.operations.furthest_point_sample(xyz, m)
    e1.record(); torch.cuda.synchronize()
    print("fps", n, m, round(e0.elapsed_time(e1) / 3, 3), "ms", round(e0.elapsed_time(e1) / 3 / m * 1e3, 3), "us/round")
