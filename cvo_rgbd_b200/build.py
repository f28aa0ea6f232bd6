"""Builds libcvo_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "cvo_api.cu")
DEPS = [SRC, os.path.join(_HERE, "..", "include", "cvo_b200.h")] + sorted(
    os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc")) if f.endswith(".cuh"))
OUT = os.path.join(_HERE, "libcvo_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build_library(force=False, verbose=False, out=None, defines=()):
    """out / defines: tuning variants (scripts/build_variants.py); the product is the default build."""
    out = out or OUT
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc]
    if os.path.exists("/usr/bin/g++"):  # the image exports CXX=/opt/gcc/bin/g++; use the distro host compiler
        cmd += ["-ccbin", "/usr/bin/g++"]
    cmd += NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + ["-o", out, SRC]
    subprocess.run(cmd, check=True)
    return out


def _build_example(name, extra_deps=(), cuda_runtime=False):
    root = os.path.dirname(_HERE)
    src = os.path.join(root, "examples", name + ".cpp")
    out = os.path.join(root, "examples", name)
    deps = [src, os.path.join(root, "include", "cvo_b200_frontend.hpp"), os.path.join(root, "include", "cvo_b200.h"),
            OUT] + [os.path.join(root, "include", d) for d in extra_deps]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-o", out, src, "-L" + _HERE, "-lcvo_b200", "-Wl,-rpath,$ORIGIN/../cvo_rgbd_b200"]
    if cuda_runtime:  # the example itself asks the runtime how many devices there are
        cmd += ["-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True)
    return out


def build_frontend_example():
    """Compiles examples/frontend_example.cpp (C++ host classes above the C ABI) against the in-tree library."""
    return _build_example("frontend_example")


def build_sequence_driver():
    """Compiles examples/cvo_sequence.cpp: the reference's sequence driver (src/cvo_main.cpp) over PCD files."""
    return _build_example("cvo_sequence", ("cvo_b200_io.hpp",))


def build_multi_gpu_example():
    """Compiles examples/cvo_batch_multi_gpu.cpp: one context per GPU + cvo_b200_align_multi (BASELINE config 4)."""
    return _build_example("cvo_batch_multi_gpu", cuda_runtime=True)


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
