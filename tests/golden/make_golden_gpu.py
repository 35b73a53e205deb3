"""Mint tests/golden/ref_cuda_*.npz from the REFERENCE'S OWN CUDA kernels (oracle/_ref, built by
oracle/build_ref.py from /root/reference/sampling and /root/reference/losses) on a B200.

Run on the GPU box:   gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/ and commit.  These fixtures pin the CPU oracle
(oracle/oracle_c.c): tests/test_cpu_oracle_golden.py replays them without a GPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def cloud(rng, b, n, snap=False):
    x = (rng.random((b, n, 3), dtype=np.float32) * 2 - 1)
    if snap:
        x = np.round(x * 8) / 8
    return x.astype(np.float32)


def main(out):
    os.makedirs(out, exist_ok=True)
    rs, rl = build_ref.load()
    assert rs is not None, "oracle/_ref missing"
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(20240229)
    fps_cases = {}
    for name, (b, n, m, snap) in {"small": (2, 312, 64, False), "ties": (2, 700, 350, True), "n513": (1, 513, 100, False),
                                  "n2496": (1, 2496, 40, False), "b32": (32, 100, 20, False),
                                  "n6240": (1, 6240, 300, False)}.items():
        x = cloud(rng, b, n, snap)
        xt = torch.from_numpy(x).to(dev)
        idx = torch.empty(b, m, dtype=torch.int32, device=dev)
        temp = torch.full((b, n), 1e10, device=dev)
        rs.furthest_sampling(b, n, m, xt, temp, idx)
        torch.cuda.synchronize()
        fps_cases[f"{name}_xyz"] = x
        fps_cases[f"{name}_idx"] = idx.cpu().numpy()
        fps_cases[f"{name}_temp"] = temp.cpu().numpy()
    np.savez_compressed(os.path.join(out, "ref_cuda_fps.npz"), **fps_cases)

    nmd = {}
    for name, (b, n, m, snap) in {"train": (4, 624, 624, False), "ragged": (2, 100, 1300, False),
                                  "ties": (1, 300, 300, True)}.items():
        x1, x2 = cloud(rng, b, n, snap), cloud(rng, b, m, snap)
        t1, t2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
        d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, dtype=torch.int32, device=dev)
        d2 = torch.empty(b, m, device=dev); i2 = torch.empty(b, m, dtype=torch.int32, device=dev)
        rl.nmdistance_forward(t1, t2, d1, d2, i1, i2)
        g1 = rng.standard_normal((b, n)).astype(np.float32); g2 = rng.standard_normal((b, m)).astype(np.float32)
        gx1 = torch.zeros_like(t1); gx2 = torch.zeros_like(t2)
        rl.nmdistance_backward(t1, t2, gx1, gx2, torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev), i1, i2)
        torch.cuda.synchronize()
        for k, v in dict(xyz1=x1, xyz2=x2, dist1=d1, idx1=i1, dist2=d2, idx2=i2, g1=g1, g2=g2, gx1=gx1, gx2=gx2).items():
            nmd[f"{name}_{k}"] = v.cpu().numpy() if torch.is_tensor(v) else v
    np.savez_compressed(os.path.join(out, "ref_cuda_nmdistance.npz"), **nmd)

    ga = {}
    pts = rng.standard_normal((2, 5, 200)).astype(np.float32)
    idx = rng.integers(0, 200, (2, 77)).astype(np.int32)
    pt, it = torch.from_numpy(pts).to(dev), torch.from_numpy(idx).to(dev)
    o = torch.empty(2, 5, 77, device=dev)
    rs.gather_forward(2, 5, 200, 77, pt, it, o)
    go = rng.standard_normal((2, 5, 77)).astype(np.float32)
    gp = torch.zeros(2, 5, 200, device=dev)
    rs.gather_backward(2, 5, 200, 77, torch.from_numpy(go).to(dev), it, gp)
    torch.cuda.synchronize()
    ga.update(points=pts, idx=idx, out=o.cpu().numpy(), grad_out=go, grad_points=gp.cpu().numpy())
    np.savez_compressed(os.path.join(out, "ref_cuda_gather.npz"), **ga)
    print("golden written to", out, {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
