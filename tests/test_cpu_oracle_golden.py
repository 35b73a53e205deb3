"""Pin the CPU oracle (oracle/oracle_c.c) against golden vectors produced by the REFERENCE'S OWN CUDA kernels
on a B200 (tests/golden/ref_cuda_*.npz, minted by tests/golden/make_golden_gpu.py from oracle/_ref).
Runs without a GPU."""
import os

import numpy as np
import pytest

from oracle import c_oracle
from tests.util import bits_equal

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(G, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not minted yet")
    return np.load(path)


@pytest.mark.parametrize("case", ["small", "ties", "n513", "n2496", "b32", "n6240"])
def test_fps_oracle_equals_reference_kernel(case):
    z = _load("ref_cuda_fps.npz")
    xyz, idx, temp = z[f"{case}_xyz"], z[f"{case}_idx"], z[f"{case}_temp"]
    t = np.full(temp.shape, 1e10, np.float32)
    got = c_oracle.fps(xyz, idx.shape[1], temp=t, legacy_temp_rows=True)
    assert np.array_equal(got, idx)
    assert bits_equal(t, temp)


@pytest.mark.parametrize("case", ["train", "ragged", "ties"])
def test_nmdistance_oracle_equals_reference_kernel(case):
    z = _load("ref_cuda_nmdistance.npz")
    d1, i1, d2, i2 = c_oracle.nmdist_fwd(z[f"{case}_xyz1"], z[f"{case}_xyz2"])
    assert bits_equal(d1, z[f"{case}_dist1"]) and bits_equal(d2, z[f"{case}_dist2"])
    assert np.array_equal(i1, z[f"{case}_idx1"]) and np.array_equal(i2, z[f"{case}_idx2"])
    gx1, gx2 = c_oracle.nmdist_bwd(z[f"{case}_xyz1"], z[f"{case}_xyz2"], z[f"{case}_g1"], z[f"{case}_g2"], i1, i2)
    np.testing.assert_allclose(gx1, z[f"{case}_gx1"], rtol=1e-5, atol=1e-6)  # atomics: order unspecified
    np.testing.assert_allclose(gx2, z[f"{case}_gx2"], rtol=1e-5, atol=1e-6)


def test_gather_oracle_equals_reference_kernel():
    z = _load("ref_cuda_gather.npz")
    assert bits_equal(c_oracle.gather_fwd(z["points"], z["idx"]), z["out"])
    np.testing.assert_allclose(c_oracle.gather_bwd(z["grad_out"], z["idx"], z["points"].shape[2]), z["grad_points"],
                               rtol=1e-5, atol=1e-6)


def test_opt_n_threads_matches_reference_formula():
    # cuda_utils.h:9-14 uses log(double)/log(2.0) truncated; must equal floor(log2 n) for every n we can meet
    for n in list(range(1, 5000)) + [2 ** i + d for i in range(12, 19) for d in (-1, 0, 1)]:
        want = min(1 << (n.bit_length() - 1), 512)
        assert c_oracle.opt_n_threads(n) == want, n
