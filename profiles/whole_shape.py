"""BASELINE config 5: whole-shape inference, num_shape_point=5000 -> 48 patches -> 16x -> FPS 239616 -> 80000
(main.py:333-389 without file IO).  Prints device time per stage."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_net
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
net.load_state_dict(ref_net.make_params(4, seed=1), strict=True)
net = net.to(dev).eval()
g = torch.Generator().manual_seed(0)
pc = ref_net.normalize_point_batch(torch.rand(1, 3, 5000, generator=g))[0].to(dev)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

t_all, out = timed(lambda: pu3.pipeline.upsample_shape(net, pc, num_point=312, patch_num_ratio=3, up_ratio=16))
t_pred, (patches, up) = timed(lambda: pu3.pipeline.pc_prediction(net, pc, 312, 3, 16))
pred = up.permute(1, 0, 2).reshape(1, 3, -1).contiguous()
t_fps, _ = timed(lambda: pu3.operations.furthest_point_sample(pred, 80000))
print(f"whole shape 5000 -> {tuple(out.shape)}: total {t_all:.1f} ms  (48 patches x16: {t_pred:.1f} ms, final FPS {pred.shape[2]} -> 80000: {t_fps:.1f} ms)")
print(f"patches/s through the shape pipeline: {48 / (t_pred / 1e3):.0f}")
