"""A/B of the exact bounding-sphere pre-filter of knn_thread_kernel at the skip connection's level-3 / level-4 shapes."""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pu3 = importlib.import_module("3pu_pytorch_b200")
dev = torch.device("cuda:0")
lib = ctypes.CDLL(pu3._lib.LIB_PATH)
g = torch.Generator().manual_seed(0)
B = 32
for n_prev_unique, n_cur, P_prev in ((624, 1248, 10), (1248, 2496, 20)):
    def sphere(n):
        v = torch.randn(B, 3, n, generator=g); return v / v.norm(dim=1, keepdim=True)
    prev_u = sphere(n_prev_unique).to(dev); cur = sphere(n_cur).to(dev)
    # previous cloud = P_prev tiles of 312 side by side (5x duplicates); tiles of the current cloud = 312-NN of FPS seeds
    _, seeds_p = pu3.operations.furthest_point_sample(prev_u, P_prev)
    prev_tiles, _, _ = pu3.operations.group_knn(312, seeds_p, prev_u, unique=False)
    prev = prev_tiles.permute(0, 1, 2, 3).reshape(B, 3, P_prev * 312).contiguous()
    P = int(n_cur / 312 * 5)
    _, seeds = pu3.operations.furthest_point_sample(cur, P)
    tiles, _, _ = pu3.operations.group_knn(312, seeds, cur, unique=False)
    q = tiles.permute(0, 2, 1, 3).reshape(B * P, 3, 312).contiguous()
    owner = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(P)
    rg = pu3.operations.Ragged(owner, owner, B)
    res = {}
    for off in (1, 0):
        lib.pu3_knn_no_prefilter(off)
        for _ in range(2):
            _, idx, _ = pu3.operations._knn_raw(5, q, prev, True, None, want_knn=False, want_dist=False, ragged=rg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            _, idx, _ = pu3.operations._knn_raw(5, q, prev, True, None, want_knn=False, want_dist=False, ragged=rg)
        e1.record(); torch.cuda.synchronize()
        res[off] = (e0.elapsed_time(e1) / 5, idx.clone())
    lib.pu3_knn_no_prefilter(0)
    print(f"prev {prev.shape[2]} (unique {n_prev_unique}), {B*P} tiles: unfiltered {res[1][0]:.3f} ms, filtered {res[0][0]:.3f} ms, equal {torch.equal(res[0][1], res[1][1])}")
