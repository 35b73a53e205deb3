"""CPU checks of the formats / input-side helpers either side of the hot path:
checkpoint dictionary (utils/pytorch_utils.py:7-51), .xyz / PLY files (utils/pc_utils.py:223-285), and the
tensor versions of the numpy augmentation helpers (utils/pc_utils.py:11-79) against numpy restatements."""
import importlib
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def pu3():
    return importlib.import_module("3pu_pytorch_b200")


def test_checkpoint_round_trip_and_reference_layout(pu3, tmp_path):
    net = torch.nn.Sequential(torch.nn.Conv1d(3, 8, 1), torch.nn.ReLU(), torch.nn.Conv1d(8, 3, 1))
    path = pu3.formats.save_network(net, str(tmp_path), "final", "poisson", step=1234)
    assert os.path.basename(path) == "final_poisson.pth"
    blob = torch.load(path, weights_only=False)
    assert list(blob.keys()) == ["states", "step"] and set(blob["states"]) == set(net.state_dict())   # the reference's layout
    other = torch.nn.Sequential(torch.nn.Conv1d(3, 8, 1), torch.nn.ReLU(), torch.nn.Conv1d(8, 3, 1))
    assert pu3.formats.load_network(other, path) == 1234
    for a, b in zip(net.parameters(), other.parameters()):
        assert torch.equal(a, b)
    # extra keys in the file are dropped (pytorch_utils.py:34-39); a missing key returns step 0 (:43-45)
    blob["states"]["not.in.model"] = torch.zeros(1)
    torch.save(blob, path)
    assert pu3.formats.load_network(other, path) == 1234
    del blob["states"]["0.weight"]
    torch.save(blob, path)
    assert pu3.formats.load_network(other, path) == 0


def test_net_state_dict_loads_a_reference_style_checkpoint(pu3, tmp_path):
    from oracle import ref_net
    params = ref_net.make_params(4, seed=3)           # the reference's state_dict keys and shapes
    torch.save({"states": params, "step": 77}, str(tmp_path / "final_scan.pth"))
    net = pu3.Net(max_up_ratio=16, step_ratio=2, knn=32, growth_rate=12, dense_n=3, fm_knn=5)
    assert pu3.formats.load_network(net, str(tmp_path / "final_scan.pth")) == 77
    sd = net.state_dict()
    assert set(sd) == set(params) and all(torch.equal(sd[k], params[k]) for k in params)


def test_ply_and_xyz_round_trip(pu3, tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.random((100, 3)).astype(np.float32)
    nrm = rng.random((100, 3)).astype(np.float32)
    col = rng.random((100, 3))
    f = str(tmp_path / "sub" / "cloud.ply")
    pu3.formats.save_ply(pts, f, colors=col, normals=nrm)
    raw = open(f, "rb").read()
    header = raw[:raw.index(b"end_header\n") + len(b"end_header\n")].decode()
    assert header == ("ply\nformat binary_little_endian 1.0\nelement vertex 100\nproperty float x\nproperty float y\n"
                      "property float z\nproperty float nx\nproperty float ny\nproperty float nz\nproperty uchar red\n"
                      "property uchar green\nproperty uchar blue\nend_header\n")    # what plyfile writes for this table
    back = pu3.formats.read_ply(f)
    assert back.shape == (100, 9)
    assert np.array_equal(back[:, :3].astype(np.float32), pts) and np.array_equal(back[:, 3:6].astype(np.float32), nrm)
    assert np.array_equal(back[:, 6:], (col * 255).astype(np.uint8))
    assert np.array_equal(pu3.formats.load(f), pts) and pu3.formats.load(f, count=10).shape == (10, 3)
    # ascii PLY with a face element after the vertices
    a = str(tmp_path / "a.ply")
    with open(a, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment x\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 1 2\n3 4 5.5\n3 0 1 1\n")
    assert np.array_equal(pu3.formats.read_ply(a), np.array([[0, 1, 2], [3, 4, 5.5]]))
    # .xyz text
    x = str(tmp_path / "out" / "pred.xyz")
    pu3.formats.save_xyz(torch.from_numpy(pts).t().contiguous().t(), x)
    assert np.allclose(pu3.formats.load(x), pts, atol=1e-6)
    padded = pu3.formats.load(x, count=130, generator=np.random.default_rng(1))
    assert padded.shape == (130, 3) and np.allclose(padded[:100], pts, atol=1e-6)
    assert all(any(np.allclose(p, q, atol=1e-6) for q in pts) for p in padded[100:105])   # padding repeats real points


def _np_rotation(angles):   # utils/pc_utils.py:54-64 restated
    Rx = np.array([[1, 0, 0], [0, np.cos(angles[0]), -np.sin(angles[0])], [0, np.sin(angles[0]), np.cos(angles[0])]])
    Ry = np.array([[np.cos(angles[1]), 0, np.sin(angles[1])], [0, 1, 0], [-np.sin(angles[1]), 0, np.cos(angles[1])]])
    Rz = np.array([[np.cos(angles[2]), -np.sin(angles[2]), 0], [np.sin(angles[2]), np.cos(angles[2]), 0], [0, 0, 1]])
    return Rz @ (Ry @ Rx)


def test_augmentation_helpers_match_numpy_restatement(pu3):
    P = pu3.patches
    g = torch.Generator().manual_seed(0)
    inp, lab = torch.rand(4, 312, 3, generator=g, dtype=torch.float64), torch.rand(4, 624, 3, generator=g, dtype=torch.float64)
    angles = torch.rand(4, 3, generator=g, dtype=torch.float64) * 2 * np.pi
    got_in, got_lab = P.augment(inp, lab, angles=angles)
    # numpy: normalise by the label's centroid / radius (data.py:153-155), then rotate (pc_utils.py:66,73)
    lab_np, inp_np = lab.numpy().copy(), inp.numpy().copy()
    c = lab_np.mean(axis=1, keepdims=True)
    lab_np = lab_np - c
    r = np.amax(np.sqrt(np.sum(lab_np ** 2, axis=-1, keepdims=True)), axis=1, keepdims=True)
    lab_np, inp_np = lab_np / r, (inp_np - c) / r
    for k in range(4):
        R = _np_rotation(angles[k].numpy())
        inp_np[k], lab_np[k] = inp_np[k] @ R, lab_np[k] @ R
    assert np.allclose(got_in.numpy(), inp_np, atol=1e-12) and np.allclose(got_lab.numpy(), lab_np, atol=1e-12)
    assert abs(float(got_lab.norm(dim=-1).max()) - 1.0) < 1e-12          # label radius 1 survives the rotation
    j = P.jitter_perturbation_point_cloud(inp, sigma=0.005, clip=0.02, generator=g)
    assert float((j - inp).abs().max()) <= 0.02 + 1e-12 and float((j - inp).abs().max()) > 0
    n2, c2, f2 = P.normalize_point_cloud(inp[0])                           # 2-D input branch (pc_utils.py:16-17)
    assert n2.shape == (312, 3) and c2.shape == (1, 3) and f2.shape == (1, 1)


def test_device_side_tile_count_equals_the_reference_host_arithmetic():
    """Net.static_tiles computes the per-request tile count int(N'/k*5) (upsampler.py:76) on the device as
    (counts.double() / k * 5).to(int32); it must equal Python's float arithmetic for every count that can occur."""
    for k in (312, 100, 77):
        counts = torch.arange(k, 25001, dtype=torch.int64)
        dev_side = (counts.double() / k * 5).to(torch.int32)
        host_side = torch.tensor([int(c / k * 5) for c in counts.tolist()], dtype=torch.int32)
        assert torch.equal(dev_side, host_side), k
    # the static slot count bounds every per-request count (N' <= N)
    for n in (624, 1248, 2496, 5000):
        assert int(n / 312 * 5) >= int((n - 1) / 312 * 5)


def test_model_resumes_from_opt_ckpt_without_extra_arguments(pu3, tmp_path):
    """model.py:25-28: Model(net, phase, opt) always loads opt.ckpt (it was silently ignored)."""
    import types
    from oracle import ref_net
    params = ref_net.make_params(1, seed=5)
    torch.save({"states": params, "step": "4321"}, str(tmp_path / "final_poisson.pth"))   # the reference stores str(step)
    net = pu3.Net(max_up_ratio=2, step_ratio=2, knn=16, growth_rate=12, dense_n=3, fm_knn=5)
    opt = types.SimpleNamespace(ckpt=str(tmp_path / "final_poisson.pth"), lr_init=5e-4)
    model = pu3.Model(net, "test", opt)
    assert model.step == 4321
    sd = net.state_dict()
    assert all(torch.equal(sd[k], params[k]) for k in params)
    assert pu3.Model(net, "test", types.SimpleNamespace(ckpt=None)).step == 0
