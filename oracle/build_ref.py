"""Compile the REFERENCE's own CUDA extensions for sm_100a into oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE ONLY.  The result (ref_sampling*.so, ref_losses*.so) is the strongest
oracle available for the native part of the path: the reference's kernels themselves, run
on the B200 next to ours (tests/test_gpu_vs_reference_cuda.py, tests/golden/make_golden_gpu.py,
and the "kernel to beat" table: profiles/kernels_to_beat.py, embedded by bench.py as `kernels_to_beat`).

Sources are read where they lie under /root/reference and are NOT copied into the repository.
losses/* compiles unmodified.  sampling/* needs the API-drift edits listed in SURVEY.md
section 8(c) (torch 1.0 -> 2.11: Tensor.type() as dispatch argument, AT_CHECK, one missing include);
they are applied by sed into a temporary directory outside the repo, compiled from there and the
temporary copy is deleted.  No kernel code is touched.

Runs only where /root/reference exists (this container).  nvcc cross-compiles without a GPU.
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("PU3_REFERENCE_ROOT", "/root/reference")

NVCC_FLAGS = ["-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-w"]  # the reference builds with nvcc -O2 (sampling/setup.py:10)
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-w"]


def _patched_copy(src, dst):
    text = open(src).read()
    name = os.path.basename(src)
    if name == "sampling_cuda.cu":
        text = "#include <ATen/cuda/CUDAContext.h>\n" + text
        text = text.replace("points.type()", "points.scalar_type()")
        text = text.replace("grad_out.type()", "grad_out.scalar_type()")
        text = text.replace("xyz.type()", "xyz.scalar_type()")
    if name == "sampling.cpp":
        text = text.replace("AT_CHECK(", "TORCH_CHECK(")
        text = re.sub(r"x\.type\(\)\.is_cuda\(\)", "x.is_cuda()", text)
        text = text.replace("query.type().is_cuda()", "query.is_cuda()")
    open(dst, "w").write(text)


def _torch_flags():
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths() + [sysconfig.get_paths()["include"]]
    libdir = ce.library_paths()
    return inc, libdir


def _build(modname, srcdir, files, patch):
    import torch  # noqa: F401
    inc, libdirs = _torch_flags()
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(OUT, modname + suffix)
    tmp = tempfile.mkdtemp(prefix="pu3_refbuild_")
    try:
        objs = []
        for f in os.listdir(srcdir):
            if f.endswith(".h"):
                shutil.copy(os.path.join(srcdir, f), os.path.join(tmp, f))
        for f in files:
            src = os.path.join(srcdir, f)
            work = os.path.join(tmp, f)
            if patch:
                _patched_copy(src, work)
            else:
                work = src  # compiled in place, untouched
            obj = os.path.join(tmp, f + ".o")
            defs = [f"-DTORCH_EXTENSION_NAME={modname}", "-DTORCH_API_INCLUDE_EXTENSION_H",
                    "-D_GLIBCXX_USE_CXX11_ABI=1"]
            incs = [f"-I{p}" for p in inc] + [f"-I{tmp}", f"-I{srcdir}"]
            if f.endswith(".cu"):
                cmd = ["nvcc"] + NVCC_FLAGS + defs + incs + ["-c", work, "-o", obj]
            else:
                cmd = ["g++"] + CXX_FLAGS + defs + incs + ["-c", work, "-o", obj]
            subprocess.check_call(cmd)
            objs.append(obj)
        link = ["g++", "-shared", "-o", target] + objs + [f"-L{p}" for p in libdirs] + \
               ["-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                "-ltorch_python", "-lcudart"] + [f"-Wl,-rpath,{p}" for p in libdirs]
        subprocess.check_call(link)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return target


def build(force=False):
    """Build both extensions; returns the list of .so paths, or [] when /root/reference is absent."""
    if not os.path.isdir(os.path.join(REF_ROOT, "sampling")):
        return sorted(glob.glob(os.path.join(OUT, "*.so")))
    os.makedirs(OUT, exist_ok=True)
    done = []
    for modname, sub, files, patch in (
            ("ref_losses", "losses", ["nmdistance.cpp", "nmdistance_cuda.cu"], False),
            ("ref_sampling", "sampling", ["sampling.cpp", "sampling_cuda.cu"], True)):
        have = glob.glob(os.path.join(OUT, modname + "*.so"))
        if have and not force:
            done += have
            continue
        done.append(_build(modname, os.path.join(REF_ROOT, sub), files, patch))
    stage_python(force)
    return done


PY_ARCHIVE = os.path.join(OUT, "reference_driver.tar.gz")
# the reference's DRIVER side, which must keep running unchanged on top of the drop-in modules: main.py, its own Model wrapper,
# the data set and the numpy / IO helpers.  NOT included: network/, sampling/, losses/ -- exactly what this repository replaces
# (3pu_pytorch_b200/shim provides those import names).
PY_FILES = ["main.py", "model.py", "data.py", "utils/__init__.py", "utils/pc_utils.py", "utils/pytorch_utils.py",
            "utils/interactive_visualizer.py", "misc/__init__.py", "misc/logger.py"]


def stage_python(force=False):
    """Pack the reference's driver files, byte for byte, into ONE git-ignored archive under oracle/_ref (it travels to the GPU box,
    where /root/reference does not exist; no reference source file is ever placed in the tree).  Returns the archive or None."""
    import tarfile
    if not os.path.isfile(os.path.join(REF_ROOT, "main.py")):
        return PY_ARCHIVE if os.path.isfile(PY_ARCHIVE) else None
    if os.path.isfile(PY_ARCHIVE) and not force:
        return PY_ARCHIVE
    os.makedirs(OUT, exist_ok=True)
    with tarfile.open(PY_ARCHIVE, "w:gz") as tar:
        for rel in PY_FILES:
            src = os.path.join(REF_ROOT, rel)
            if not os.path.isfile(src):
                if rel.endswith("__init__.py"):      # the reference uses namespace packages where it has no __init__
                    continue
                raise FileNotFoundError(src)
            tar.add(src, arcname=rel)
    shutil.rmtree(os.path.join(OUT, "py"), ignore_errors=True)     # an earlier layout staged loose files
    return PY_ARCHIVE


def unpack_python(dest):
    """Extract the staged driver files into `dest` (a test's temporary directory); returns dest or None when not staged."""
    import tarfile
    arc = stage_python()
    if arc is None:
        return None
    with tarfile.open(arc, "r:gz") as tar:
        tar.extractall(dest, filter="data")
    return dest


def load():
    """Import (ref_sampling, ref_losses) from oracle/_ref, or (None, None) if not built."""
    import importlib
    import torch  # noqa: F401  (the extensions link against libtorch)
    if not glob.glob(os.path.join(OUT, "ref_sampling*.so")) or not glob.glob(os.path.join(OUT, "ref_losses*.so")):
        return None, None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    return importlib.import_module("ref_sampling"), importlib.import_module("ref_losses")


if __name__ == "__main__":
    print("\n".join(build(force="--force" in sys.argv)))
