"""Where does the whole-shape pipeline (main.py:214-246,346-380) diverge from the oracle?  Stage by stage, per patch."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_net
from tests.util import cloud_match_fraction
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
for seed in (11, 1):
    params = ref_net.make_params(4, seed=seed)
    net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
    net.load_state_dict(params, strict=True); net = net.to(dev).eval()
    g = torch.Generator().manual_seed(77)
    pts = torch.rand(624, 3, generator=g).numpy().astype(np.float32) * np.float32(2.0) + np.float32(0.5)
    data = pts[np.newaxis]
    c = np.mean(data, axis=1, keepdims=True); data = data - c
    far = np.amax(np.sqrt(np.sum(data ** 2, axis=-1, keepdims=True)), axis=1, keepdims=True); data = data / far
    pc = torch.from_numpy(data).transpose(2, 1).contiguous()
    with torch.no_grad():
        P = int(624 / 312 * 3)
        i_g, s_g = pu3.operations.furthest_point_sample(pc.to(dev), P)
        i_o, s_o = ref_net.furthest_point_sample(pc, P)
        print(f"seed {seed}: FPS seeds equal: {torch.equal(i_g.cpu(), i_o)}")
        p_g, _, _ = pu3.operations.group_knn(312, s_g, pc.to(dev))
        p_o, _, _ = ref_net.group_knn(312, s_o, pc, unique=True)
        print("  patches max diff", float((p_g.cpu() - p_o).abs().max()))
        for k in range(P):
            pg, cg, rg = pu3.operations.normalize_point_batch(p_g[:, :, k, :])
            po, co, ro = ref_net.normalize_point_batch(p_o[:, :, k, :])
            for ratio in (2, 4, 8, 16):
                ug = net(pg, ratio=ratio).cpu()
                uo = ref_net.net_forward(params, po, ratio=ratio, max_up_ratio=16, knn=32)
                print(f"  patch {k} ratio {ratio}: match(1e-4) = {cloud_match_fraction(ug[0], uo[0], tol=1e-4):.4f}  max|coord| {float(uo.abs().max()):.2f}")
