"""On-disk formats either side of the path: the reference's checkpoint dictionary
(utils/pytorch_utils.py:7-51) and its point-cloud files (.xyz text and binary little-endian PLY,
utils/pc_utils.py:223-285).  plyfile is not in this image, so PLY is read and written directly; files written
here open in the reference's plyfile-based reader and vice versa (same header, same property order)."""
import os
from collections import OrderedDict

import numpy as np
import torch


# ---- checkpoints -------------------------------------------------------------------------------------------------
def save_network(net, directory, network_label, epoch_label=None, **kwargs):
    """utils/pytorch_utils.py:7-15: {'states': state_dict on CPU, **kwargs} -> <label>_<epoch>.pth.  The module is
    not moved (the reference does net.cpu() ... net.cuda(), which breaks flat-buffer parameter views)."""
    save_path = os.path.join(directory, "_".join((network_label, str(epoch_label))) + ".pth")
    merged = OrderedDict()
    merged["states"] = OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items())
    for k, v in kwargs.items():
        merged[k] = v
    os.makedirs(directory, exist_ok=True)
    torch.save(merged, save_path)
    return save_path


def load_network(net, path, map_location="cpu"):
    """utils/pytorch_utils.py:18-51: loads 'states', dropping keys the model does not have; returns the stored step
    (0 if absent or if a key the model needs is missing, like the reference)."""
    if path.endswith("pth"):
        loaded = torch.load(path, map_location=map_location, weights_only=False)
    else:
        loaded = np.load(path, allow_pickle=True).item()
    network = net.module if isinstance(net, torch.nn.DataParallel) else net
    own = network.state_dict()
    for k in set(loaded["states"].keys()) - set(own.keys()):
        del loaded["states"][k]
    try:
        network.load_state_dict(loaded["states"])
    except (KeyError, RuntimeError) as e:     # torch >= 1.x raises RuntimeError for missing keys
        print(e)
        return 0
    return loaded.get("step", 0)


# ---- point clouds ------------------------------------------------------------------------------------------------
_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
              "float32": "f4", "float64": "f8"}


def read_ply(filename, count=None):
    """Vertex table of a PLY file (ascii or binary) as a float (N, n_properties) array, first `count` rows."""
    with open(filename, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{filename}: not a PLY file")
        fmt, props, nvert, in_vertex = None, [], 0, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{filename}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    nvert = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{filename}: list property on vertices is not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=nvert, ndmin=2)[:, :len(props)]
        else:
            order = "<" if fmt == "binary_little_endian" else ">"
            rec = np.fromfile(f, dtype=np.dtype([(n, order + t) for n, t in props]), count=nvert)
            data = np.stack([rec[n].astype(np.float64) for n, _ in props], axis=1)
    return data[:count] if count is not None else data


def load(filename, count=None, generator=None):
    """utils/pc_utils.py:223-243: .ply -> xyz of the vertices; anything else -> np.loadtxt.  With `count`, short
    clouds are padded with randomly repeated points; long ones are cut to the first `count` rows (the reference
    calls an FPS-based downsample there; use operations.furthest_point_sample on the GPU for that)."""
    if filename.endswith(".ply"):
        return read_ply(filename, count)[:, :3].astype(np.float32)
    points = np.loadtxt(filename, ndmin=2).astype(np.float32)
    if count is not None and count > points.shape[0]:
        rng = generator or np.random.default_rng()
        extra = points[rng.choice(points.shape[0], count - points.shape[0])]
        points = np.concatenate([points, extra], axis=0)
    elif count is not None:
        points = points[:count]
    return points


def save_xyz(points, filename):
    """main.py:383 (np.savetxt of the (N,3) prediction)."""
    points = points.detach().cpu().numpy() if torch.is_tensor(points) else np.asarray(points)
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    np.savetxt(filename, points.reshape(-1, points.shape[-1]), fmt="%.6f")


def save_ply(points, filename, colors=None, normals=None):
    """utils/pc_utils.py:246-285: binary little-endian PLY, properties x y z [nx ny nz] [red green blue [alpha]]."""
    points = points.detach().cpu().numpy() if torch.is_tensor(points) else np.asarray(points)
    points = points.reshape(-1, 3).astype("<f4")
    n = points.shape[0]
    desc = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    cols = {"x": points[:, 0], "y": points[:, 1], "z": points[:, 2]}
    if normals is not None:
        normals = np.asarray(normals, dtype="<f4").reshape(-1, 3)
        assert normals.shape[0] == n
        for i, name in enumerate(("nx", "ny", "nz")):
            desc.append((name, "<f4")); cols[name] = normals[:, i]
    if colors is not None:
        colors = np.asarray(colors)
        assert colors.shape[0] == n
        if colors.max() <= 1:
            colors = colors * 255
        for i, name in enumerate(("red", "green", "blue", "alpha")[:colors.shape[1]]):
            desc.append((name, "u1")); cols[name] = colors[:, i].astype("u1")
    rec = np.empty(n, dtype=np.dtype(desc))
    for name, _ in desc:
        rec[name] = cols[name]
    names = {"<f4": "float", "u1": "uchar"}
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    header += [f"property {names[t]} {name}" for name, t in desc] + ["end_header"]
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    with open(filename, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        rec.tofile(f)
