"""The CMake build (CMakeLists.txt; promised in BASELINE.md section 2 / SURVEY.md 8d, reference side:
cpp/rkhs_registration/CMakeLists.txt): configures for sm_100a, builds libcvo_b200.so, the C++ examples and the oracle,
and -- where the reference tree is present -- the checkers compiled from it through -DCVO_REF_DIR.  No GPU needed:
nvcc cross-compiles; the test checks the exported C ABI and that the CMake-built oracle computes what the in-tree one does."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from cvo_rgbd_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/cpp/rkhs_registration"


@pytest.mark.skipif(shutil.which("cmake") is None or shutil.which("ninja") is None, reason="cmake / ninja not installed")
def test_cmake_builds_the_library_the_examples_and_the_oracle(tmp_path, oracle):
    bdir = str(tmp_path / "b")
    cmd = ["cmake", "-S", ROOT, "-B", bdir, "-G", "Ninja", "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc"]
    if os.path.exists("/usr/bin/g++"):  # the image's default CXX is a wrapper without OpenMP
        cmd += ["-DCMAKE_CXX_COMPILER=/usr/bin/g++", "-DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++"]
    have_ref = os.path.isdir(REF)
    if have_ref:
        cmd.append("-DCVO_REF_DIR=" + REF)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    out = subprocess.run(["cmake", "--build", bdir], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for exe in ("frontend_example", "cvo_sequence", "cvo_batch_multi_gpu"):
        assert os.access(os.path.join(bdir, exe), os.X_OK)
    # every symbol of include/cvo_b200.h, built for sm_100a
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(bdir, "libcvo_b200.so")], capture_output=True, text=True).stdout
    for name in capi.EXPORTS:
        assert " T " + name + "\n" in syms, name
    dump = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", os.path.join(bdir, "libcvo_b200.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in dump and "sm_90" not in dump
    # the CMake-built oracle is the oracle
    pr = synth.make_pair(1000, 500, 500, "cvo")
    want = oracle.evaluate(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], np.eye(3), np.zeros(3), 0.1, oracle.default_params("cvo"))
    lib = C.CDLL(os.path.join(bdir, "libcvo_oracle.so"))
    from oracle.cvo_oracle import EvalOut, Params, _f32, _ptr
    p = Params()
    lib.oracle_default_params_cvo(C.byref(p))
    fp = C.POINTER(C.c_float)
    lib.oracle_eval.argtypes = [fp, fp, C.c_int, fp, fp, C.c_int, fp, fp, C.c_float, C.POINTER(Params), C.POINTER(EvalOut)]
    o = EvalOut()
    a = [_f32(pr[k]) for k in ("x_pos", "x_feat", "y_pos", "y_feat")]
    R, T = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    assert lib.oracle_eval(_ptr(a[0]), _ptr(a[1]), 500, _ptr(a[2]), _ptr(a[3]), 500, _ptr(R), _ptr(T), C.c_float(0.1), C.byref(p), C.byref(o)) == 0
    assert o.nnz == want["nnz"] and o.B == want["B"] and o.E == want["E"]
    if have_ref:
        for so in ("libcvo_oracle_ref.so", "libcvo_refsrc.so"):
            assert os.path.exists(os.path.join(bdir, so)), so
        r = C.CDLL(os.path.join(bdir, "libcvo_refsrc.so"))
        r.refsrc_backend.restype = C.c_char_p
        assert b"src/cvo.cpp" in r.refsrc_backend()
