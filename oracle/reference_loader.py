"""Import the UNMODIFIED reference Python (network/*) from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  Used in this container (where /root/reference exists)
to validate oracle/ref_net.py and to mint tests/golden/ref_py.npz (tests/golden/make_golden_py.py).  Nothing
here runs on the GPU box.

The reference imports three modules that do not exist on CPU (network/operations.py:2,6;
network/model_loss.py:2): `faiss` (dead code path), `sampling` and `losses`
(CUDA-only extensions, "CPU not supported" sampling.cpp:77).  They are provided
as stub modules backed by oracle_c.c for the duration of the import only, and
the reference package is registered under a private name so that it cannot
collide with the product's own drop-in `network` / `sampling` / `losses` shims.

One run-time fix is applied, because without it no backward can run at all:
NmDistanceFunction.backward (model_loss.py:21-28) references undefined names
d_dist1/d_dist2 and the removed ctx.saved_variables; it is replaced by the same
body without those two lines and with ctx.saved_tensors.
"""
import importlib.util
import os
import sys
import types

import torch

from . import c_oracle

REF_ROOT = os.environ.get("PU3_REFERENCE_ROOT", "/root/reference")
_PKG = "_pu3_reference_network"
_cache = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "network", "upsampler.py"))


def _stub_sampling():
    m = types.ModuleType("sampling")

    def furthest_sampling(b, n, npoint, inp, temp, idx):
        out = c_oracle.fps(inp.detach().numpy(), npoint, temp=temp.numpy())
        idx.copy_(torch.from_numpy(out))
        return idx

    def gather_forward(b, c, n, npoints, points, idx, out):
        out.copy_(torch.from_numpy(c_oracle.gather_fwd(points.detach().numpy(), idx.numpy())))
        return out

    def gather_backward(b, c, n, npoints, grad_out, idx, grad_points):
        c_oracle.gather_bwd(grad_out.detach().numpy(), idx.numpy(), n, grad_points=grad_points.numpy())
        return grad_points

    m.furthest_sampling, m.gather_forward, m.gather_backward = furthest_sampling, gather_forward, gather_backward
    return m


def _stub_losses():
    m = types.ModuleType("losses")

    def nmdistance_forward(xyz1, xyz2, dist1, dist2, idx1, idx2):
        d1, i1, d2, i2 = c_oracle.nmdist_fwd(xyz1.detach().numpy(), xyz2.detach().numpy())
        dist1.copy_(torch.from_numpy(d1)); idx1.copy_(torch.from_numpy(i1))
        dist2.copy_(torch.from_numpy(d2)); idx2.copy_(torch.from_numpy(i2))
        return 1

    def nmdistance_backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
        g1, g2 = c_oracle.nmdist_bwd(xyz1.detach().numpy(), xyz2.detach().numpy(), graddist1.contiguous().numpy(),
                                     graddist2.contiguous().numpy(), idx1.numpy(), idx2.numpy())
        gradxyz1.add_(torch.from_numpy(g1)); gradxyz2.add_(torch.from_numpy(g2))
        return 1

    m.nmdistance_forward, m.nmdistance_backward = nmdistance_forward, nmdistance_backward
    return m


def load():
    """Returns a namespace with .operations .layers .upsampler .model_loss (the reference modules)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT}")
    stubs = {"faiss": types.ModuleType("faiss"), "sampling": _stub_sampling(), "losses": _stub_losses()}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        pkg_dir = os.path.join(REF_ROOT, "network")
        spec = importlib.util.spec_from_file_location(_PKG, os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[_PKG] = pkg
        spec.loader.exec_module(pkg)
        ns = types.SimpleNamespace()
        for name in ("operations", "layers", "upsampler", "model_loss"):
            setattr(ns, name, importlib.import_module(f"{_PKG}.{name}"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    losses = stubs["losses"]

    def fixed_backward(ctx, graddist1, gradNone1, graddist2, gradNone2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1 = torch.zeros_like(xyz1)
        gradxyz2 = torch.zeros_like(xyz2)
        losses.nmdistance_backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
        return gradxyz1, gradxyz2

    ns.model_loss.NmDistanceFunction.backward = staticmethod(fixed_backward)
    _cache = ns
    return ns


def build_net(P, max_up_ratio=16, knn=32, **kw):
    """Reference Net (main.py:114-115 defaults) loaded with the flat parameter dict P."""
    ref = load()
    levels = max(int(k.split(".")[1].split("_")[1]) for k in P)
    assert 2 ** levels == max_up_ratio, (levels, max_up_ratio)
    net = ref.upsampler.Net(max_up_ratio=max_up_ratio, step_ratio=2, knn=knn, growth_rate=12, dense_n=3,
                            fm_knn=5, **kw)
    missing, unexpected = net.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    assert not missing and not unexpected
    return net
