"""Speculative FPS: time and samples-per-exchange against the speculation depth, on the merged tile clouds of a real eval step."""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import ref_net
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
lib = ctypes.CDLL(pu3._lib.LIB_PATH)
lib.pu3_fps_last_exchanges.restype = ctypes.c_uint
if '--select-all' in sys.argv:
    lib.pu3_fps_set_select_all(1)
g = torch.Generator().manual_seed(0)
clouds = {}
# (a) what the network actually produces at level 4 (xavier weights: volumetric blobs) -- capture the merged cloud via the debug hook
net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
net.load_state_dict(ref_net.make_params(4, seed=1), strict=True); net = net.to(dev).eval()
x = bench.make_inputs(0).to(dev)
net.use_cuda_graph = False
orig = net._eval_level_static
grabbed = {}
def spy(level, xyz, *a, **kw):
    dbg = {}
    out = orig(level, xyz, *a, debug=dbg, **kw)
    grabbed[xyz.shape[2]] = dbg["merged_pm"].clone()
    return out
net._eval_level_static = spy
with torch.no_grad(): net(x, ratio=16)
for n_in, merged in grabbed.items(): clouds[f"net level cloud {merged.shape[1]}"] = merged
# (b) surface-like data: tiles of points on a sphere, jittered
v = torch.randn(32, 24960, 3, generator=g); clouds["sphere surface 24960"] = (v / v.norm(dim=2, keepdim=True)).to(dev).contiguous()
for name, pts in clouds.items():
    B, n, _ = pts.shape
    m = n // 5
    idx = torch.empty(B, m, dtype=torch.int32, device=dev)
    ref = None
    for depth in (1, 2, 3, 4):
        lib.pu3_fps_set_spec(depth)
        for _ in range(2):
            pu3._lib.launch("pu3_fps_f32", pts, B, n, m, pts.data_ptr(), None, idx.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pu3._lib.launch("pu3_fps_f32", pts, B, n, m, pts.data_ptr(), None, idx.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ex = lib.pu3_fps_last_exchanges()
        same = True if ref is None else bool(torch.equal(ref, idx))
        if ref is None: ref = idx.clone()
        print(f"{name}: n={n} m={m} depth {depth}: {e0.elapsed_time(e1)/3:.3f} ms, {ex} exchanges -> {(m-1)/max(ex,1):.2f} samples/exchange, {e0.elapsed_time(e1)/3*1e3/max(ex,1):.2f} us/exchange, identical {same}")
lib.pu3_fps_set_spec(4)
