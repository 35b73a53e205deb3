"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pu3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pu3_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ("pu3_fps_f32", "pu3_gather_fwd", "pu3_gather_bwd", "pu3_nmdist_fwd_f32", "pu3_nmdist_bwd_f32",
                 "pu3_group_knn_f32", "pu3_ball_query_f32"):
        assert must in syms


def test_library_exports_every_declared_symbol(pu3):
    handle = ctypes.CDLL(pu3._lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/pu3_b200.h but not exported"
    assert set(_declared_symbols()) == set(pu3._lib.SIGNATURES), "binding table and header disagree"
    assert handle.pu3_version() == pu3._lib.ABI_VERSION


def test_argument_errors_do_not_need_a_gpu(pu3):
    L = pu3._lib.lib()
    assert L.pu3_fps_f32(1, 0, 3, None, None, None, None) == -1
    assert b"empty cloud" in L.pu3_last_error()
    assert L.pu3_gather_fwd(1, 1, 4, 2, 3, 1, 1, 1, None) == -1  # elem_bytes 3
    assert L.pu3_group_knn_f32(1, 3, 4, 2, 5, 1, 1, 1, 0, 0, None, None, None, None, None, 0, None) == -1  # k > n
    assert b"greater or equal to k" in L.pu3_last_error()
    # tensor-core convolutions: what TMA cannot address is refused before anything touches the device
    fake = ctypes.c_void_p(4096)           # aligned, never dereferenced on these paths
    assert L.pu3_conv_tc_f32(2, 30, 8, 8, fake, 240, fake, None, fake, 240, 0, None) == -1
    assert b"TMA" in L.pu3_last_error()
    assert L.pu3_conv_tc_f32(2, 32, 8, 200, fake, 256, fake, None, fake, 6400, 0, None) == -1          # cout > 128
    assert L.pu3_conv_tc_prepare_f32(8, 8, fake, 4, fake, None) == -1                                  # row stride < cin
    assert L.pu3_conv_tc_project_f32(1, 32, 8, 100, 3, fake, 256, fake, None, fake, None, fake, 96, None, 0, 1, 1, None) == -1
    assert L.pu3_conv_tc_wsplit_bytes(264, 128) == 9 * 2 * 128 * 128 and L.pu3_conv_tc_wsplit_bytes(128, 24) == 4 * 2 * 64 * 128
    assert L.pu3_conv_tc_wsplit_bytes(8, 200) == 0


def test_cpu_tensors_are_refused(pu3):
    import torch
    with pytest.raises(RuntimeError):
        pu3.operations.group_knn(3, torch.rand(1, 3, 8), torch.rand(1, 3, 8))
    with pytest.raises(RuntimeError):
        pu3.operations.furthest_point_sample(torch.rand(1, 3, 8), 2)


def test_header_is_plain_c_and_cxx():
    """The drop-in boundary is a C ABI: include/pu3_b200.h must compile as C99 and as C++17 on its own (no torch, no CUDA
    headers), and a C translation unit must link against the shared library without a C++ runtime in its own code."""
    import shutil
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "pu3_b200.h")
    gcc = shutil.which("gcc") or "/opt/gcc/bin/gcc"
    gxx = shutil.which("g++") or "/opt/gcc/bin/g++"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    lib = os.path.join(root, "3pu_pytorch_b200", "lib", "libpu3_b200.so")
    if not os.path.isfile(lib):
        pytest.skip("library not built")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "probe.c")
        with open(src, "w") as f:
            f.write('#include "pu3_b200.h"\n#include <stdio.h>\n'
                    'int main(void) { printf("%d %zu\\n", pu3_version(), pu3_conv_tc_wsplit_bytes(264, 128)); return 0; }\n')
        exe = os.path.join(d, "probe")
        subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(root, "include"), src, "-o", exe, lib,
                               "-Wl,-rpath," + os.path.dirname(lib)])
        out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) == 1 and int(out[1]) == 9 * 2 * 128 * 128        # ABI version; 9 k-blocks x (hi, lo) x 128 rows x 128 B
