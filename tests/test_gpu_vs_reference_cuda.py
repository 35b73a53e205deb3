"""GPU parity against the REFERENCE'S OWN kernels: sampling/sampling_cuda.cu and losses/nmdistance_cuda.cu
compiled unmodified (API-drift edits only, oracle/build_ref.py) for sm_100a into oracle/_ref and run on the
same B200 on the same inputs.  Skipped when oracle/_ref has not been built."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from tests.util import bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    s, l = build_ref.load()
    if s is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return s, l


@pytest.mark.parametrize("b,n,m", [(1, 312, 100), (1, 624, 10), (1, 6240, 1248), (4, 2496, 40), (32, 624, 64),
                                   (1, 24960, 600), (2, 700, 700)])
def test_fps_bit_exact_with_reference_kernel(pu3, cuda, ref, b, n, m):
    rs, _ = ref
    g = torch.Generator().manual_seed(n + m)
    x = torch.rand(b, n, 3, generator=g)
    if n == 700:  # exact ties: duplicated + grid-snapped points
        x = torch.round(x * 8) / 8
    x = x.to(cuda)
    want = torch.empty(b, m, dtype=torch.int32, device=cuda)
    t_ref = torch.full((b, n), 1e10, device=cuda)
    rs.furthest_sampling(b, n, m, x, t_ref, want)
    got = torch.empty(b, m, dtype=torch.int32, device=cuda)
    t_got = torch.full((b, n), 1e10, device=cuda)
    pu3.sampling.furthest_sampling(b, n, m, x, t_got, got)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert bits_equal(t_got.cpu().numpy(), t_ref.cpu().numpy())


@pytest.mark.parametrize("b,n,m", [(32, 624, 624), (2, 600, 2500), (1, 5000, 300)])
def test_nmdistance_bit_exact_with_reference_kernel(pu3, cuda, ref, b, n, m):
    _, rl = ref
    g = torch.Generator().manual_seed(n * 3 + m)
    x1 = torch.rand(b, n, 3, generator=g).to(cuda); x2 = torch.rand(b, m, 3, generator=g).to(cuda)
    outs = []
    for mod in (rl, pu3.losses):
        d1 = torch.empty(b, n, device=cuda); i1 = torch.empty(b, n, dtype=torch.int32, device=cuda)
        d2 = torch.empty(b, m, device=cuda); i2 = torch.empty(b, m, dtype=torch.int32, device=cuda)
        mod.nmdistance_forward(x1, x2, d1, d2, i1, i2)
        outs.append((d1, i1, d2, i2))
    torch.cuda.synchronize()
    for a, c in zip(*outs):
        assert bits_equal(a.cpu().numpy(), c.cpu().numpy())
    # backward (atomics: order differs run to run in both implementations)
    g1 = torch.randn(b, n, generator=g).to(cuda); g2 = torch.randn(b, m, generator=g).to(cuda)
    grads = []
    for mod in (rl, pu3.losses):
        gx1 = torch.zeros_like(x1); gx2 = torch.zeros_like(x2)
        mod.nmdistance_backward(x1, x2, gx1, gx2, g1, g2, outs[0][1], outs[0][3])
        grads.append((gx1, gx2))
    torch.cuda.synchronize()
    for a, c in zip(*grads):
        torch.testing.assert_close(a, c, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64])
def test_gather_matches_reference_kernel(pu3, cuda, ref, dtype):
    rs, _ = ref
    b, c, n, m = 3, 24, 700, 300
    g = torch.Generator().manual_seed(1)
    pts = torch.randn(b, c, n, generator=g).to(dtype).to(cuda)
    idx = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32).to(cuda)
    a = torch.empty(b, c, m, dtype=dtype, device=cuda); d = torch.empty_like(a)
    rs.gather_forward(b, c, n, m, pts, idx, a)
    pu3.sampling.gather_forward(b, c, n, m, pts, idx, d)
    torch.cuda.synchronize()
    assert bits_equal(a.cpu().numpy(), d.cpu().numpy())
    if dtype != torch.float16:
        go = torch.randn(b, c, m, generator=g).to(dtype).to(cuda)
        ga = torch.zeros(b, c, n, dtype=dtype, device=cuda); gd = torch.zeros_like(ga)
        rs.gather_backward(b, c, n, m, go, idx, ga)
        pu3.sampling.gather_backward(b, c, n, m, go, idx, gd)
        torch.cuda.synchronize()
        torch.testing.assert_close(ga, gd, rtol=1e-5, atol=1e-6)
