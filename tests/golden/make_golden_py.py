"""Mint tests/golden/ref_py.npz from the UNMODIFIED reference Python (/root/reference/network/*, imported on CPU by
oracle/reference_loader.py: stub `faiss`, `sampling` / `losses` backed by oracle_c.c).  Run in the build container, where
/root/reference exists:

    python tests/golden/make_golden_py.py

The fixture pins oracle/ref_net.py (the restatement every GPU parity test compares against) to the reference itself:
tests/test_cpu_ref_net_golden.py replays the stored inputs through ref_net and requires bit-identical results.  Large
outputs are stored as a SHA-256 of their bytes plus a strided sample; inputs and small outputs in full.
Weights are oracle.ref_net.make_params(levels, seed) (numpy PCG64: identical on every host), checked by a digest.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader, ref_net  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.set_num_threads(1)          # one thread: the CPU convolutions / matmuls reduce in a fixed order


def digest(t):
    a = np.ascontiguousarray(t.detach().numpy() if torch.is_tensor(t) else t)
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest(), dtype=np.uint8).copy()


def sample(t, n=257):
    a = (t.detach().numpy() if torch.is_tensor(t) else t).reshape(-1)
    return a[:: max(1, a.size // n)][:n].copy()


def main():
    ref = reference_loader.load()
    out = {}
    g = torch.Generator().manual_seed(2024)

    # ---- group_knn (operations.py:151-216): duplicates, both layouts, k == n ------------------------------------
    pts = torch.rand(2, 24, 80, generator=g)
    pts[0, :, 50] = pts[0, :, 10]; pts[0, :, 51] = pts[0, :, 10]; pts[1, :, 79] = pts[1, :, 0]
    qry = pts[:, :, :30].contiguous()
    knn, idx, dist = ref.operations.group_knn(9, qry, pts, unique=True, NCHW=True)
    out.update(knn_pts=pts.numpy(), knn_qry=qry.numpy(), knn_out=knn.contiguous().numpy(), knn_idx=idx.numpy(), knn_dist=dist.numpy())
    cloud = torch.rand(1, 200, 3, generator=g)
    seeds = cloud[:, :7].contiguous()
    knn2, idx2, dist2 = ref.operations.group_knn(20, seeds, cloud, unique=False, NCHW=False)
    out.update(knn2_cloud=cloud.numpy(), knn2_out=knn2.contiguous().numpy(), knn2_idx=idx2.numpy(), knn2_dist=dist2.numpy())

    # ---- one Level (upsampler.py:272-374) and a 2-level eval forward (:107-189) ----------------------------------
    P = ref_net.make_params(2, seed=5)
    out["params_digest"] = digest(torch.cat([P[k].reshape(-1) for k in sorted(P)]))
    net = reference_loader.build_net(P, max_up_ratio=4, knn=32).eval()
    x = ref.operations.normalize_point_batch(torch.rand(1, 3, 312, generator=g))[0]
    with torch.no_grad():
        lx, lf = net.levels["level_1"](x, x, previous_level4=None)
        up = net(x, ratio=4)
    out.update(level_in=x.numpy(), level_xyz=lx.numpy(), level_feat_sha=digest(lf), level_feat_sample=sample(lf),
               eval4_out=up.numpy())

    # ---- train-mode forward (zoom patches, fixed RNG), Chamfer loss, backward (model.py:53-77 without the optimizer) --
    netT = reference_loader.build_net(P, max_up_ratio=4, knn=32).train()
    xt = torch.rand(2, 3, 312, generator=g)
    gt = torch.rand(2, 3, 1248, generator=g)
    torch.manual_seed(77)                                      # the in-forward torch.randint (upsampler.py:55)
    pred, gt_patch = netT(xt, ratio=4, gt=gt)
    loss = ref.model_loss.ChamferLoss()(pred, gt_patch)
    loss.backward()
    out.update(train_x=xt.numpy(), train_gt=gt.numpy(), train_pred=pred.detach().numpy(), train_gt_patch=gt_patch.numpy(),
               train_loss=np.array([loss.item()], dtype=np.float64),
               train_grad_up2_sha=digest(netT.levels["level_2"].up_layer.up_layer2.conv.weight.grad),
               train_grad_up2_sample=sample(netT.levels["level_2"].up_layer.up_layer2.conv.weight.grad),
               train_grad_l1_layer0=netT.levels["level_1"].layer0.conv.weight.grad.numpy())

    # ---- ChamferLoss with the threshold branch (model_loss.py:67-77), both layouts -----------------------------------
    a, b = torch.rand(3, 100, 3, generator=g), torch.rand(3, 3, 150, generator=g)
    out["cd_a"], out["cd_b"] = a.numpy(), b.numpy()
    out["cd_plain"] = np.array([ref.model_loss.ChamferLoss()(a, b).item()], dtype=np.float64)
    out["cd_thresh"] = np.array([ref.model_loss.ChamferLoss(threshold=1.5, forward_weight=0.7)(a, b).item()], dtype=np.float64)

    import platform
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        cpu = platform.processor()
    out["meta"] = np.array([torch.__version__, cpu, str(torch.backends.mkldnn.is_available())])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_py.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
