"""GPU: the drop-in claim, end to end.  The reference's own driver -- main.py, model.py, data.py, utils/, misc/, staged byte for
byte into git-ignored oracle/_ref/py by oracle/build_ref.py -- runs UNCHANGED (`python main.py --phase test ...`,
main.py:333-389) with 3pu_pytorch_b200/shim first on PYTHONPATH providing `network`, `sampling`, `losses`, `faiss` (and import
stand-ins for plyfile / matplotlib / h5py / visdom, which this image lacks).  The PLY it writes is compared with the oracle's
walk through the same pipeline, and pins formats.save_ply against a file written by the reference's own save_ply."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import build_ref, ref_net
from tests.util import cloud_match_fraction

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def staged():
    d = build_ref.stage_python()
    if d is None:
        pytest.skip("oracle/_ref/py not staged (needs /root/reference at build time)")
    return d


def test_reference_main_py_runs_unchanged_through_the_shim(pu3, cuda, staged, tmp_path):
    n_shape, ratio = 624, 16      # 6 overlapping patches
    g = torch.Generator().manual_seed(77)
    pts = torch.rand(n_shape, 3, generator=g).numpy().astype(np.float32) * np.float32(2.0) + np.float32(0.5)
    os.makedirs(tmp_path / "shapes")
    np.savetxt(tmp_path / "shapes" / "blob.xyz", pts, fmt="%.8f")
    params = ref_net.make_params(4, seed=11)
    torch.save({"states": params, "step": "0"}, str(tmp_path / "final_synth.pth"))
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "3pu_pytorch_b200", "shim"), ROOT, env.get("PYTHONPATH", "")])
    cmd = [sys.executable, "main.py", "--phase", "test", "--id", "demo", "--ckpt", str(tmp_path / "final_synth.pth"),
           "--test_data", str(tmp_path / "shapes" / "*.xyz"), "--num_shape_point", str(n_shape), "--num_point", "312",
           "--up_ratio", str(ratio), "--result_dir", str(tmp_path / "out")]
    r = subprocess.run(cmd, cwd=staged, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    out_ply = tmp_path / "out" / "shapes" / "blob.ply"
    in_ply = tmp_path / "out" / "shapes" / "blob_input.ply"
    assert out_ply.is_file() and in_ply.is_file(), r.stdout[-1500:]
    got = pu3.formats.read_ply(str(out_ply))[:, :3].astype(np.float32)
    assert got.shape == (n_shape * ratio, 3) and np.isfinite(got).all()

    # ---- the same pipeline on the oracle (main.py:346-380 with utils/pc_utils.py:11-25 for the numpy normalisation) -------
    data = np.loadtxt(tmp_path / "shapes" / "blob.xyz").astype(np.float32)[np.newaxis]
    centroid = np.mean(data, axis=1, keepdims=True)
    data = data - centroid
    far = np.amax(np.sqrt(np.sum(data ** 2, axis=-1, keepdims=True)), axis=1, keepdims=True)
    data = data / far
    pc = torch.from_numpy(data).transpose(2, 1)                                    # 1x3xN
    with torch.no_grad():
        num_patches = int(pc.shape[2] / 312 * 3)
        _, seeds = ref_net.furthest_point_sample(pc, num_patches)
        patches, _, _ = ref_net.group_knn(312, seeds, pc, unique=True)
        ups = []
        for k in range(num_patches):
            patch, c, rad = ref_net.normalize_point_batch(patches[:, :, k, :])
            up = ref_net.net_forward(params, patch, ratio=ratio, max_up_ratio=ratio, knn=32)
            ups.append(up * rad + c)
        pred = torch.cat(ups, dim=-1)
        _, pred = ref_net.furthest_point_sample(pred, n_shape * ratio)
    want = pred.transpose(2, 1).numpy() * far + centroid
    frac = cloud_match_fraction(torch.from_numpy(got.T.copy()), torch.from_numpy(want[0].T.astype(np.float32).copy()), tol=2e-4)
    assert frac > 0.95, frac

    # ---- f-4: formats.save_ply against the file the REFERENCE's save_ply (utils/pc_utils.py:246-285) wrote -------------------
    ref_in = pu3.formats.read_ply(str(in_ply))[:, :3].astype(np.float32)
    pu3.formats.save_ply(ref_in, str(tmp_path / "mine_input.ply"))
    assert open(tmp_path / "mine_input.ply", "rb").read() == open(in_ply, "rb").read()        # header and records, byte for byte
    np.testing.assert_allclose(ref_in, (data[0] * far[0] + centroid[0]), rtol=1e-6, atol=1e-6)
