"""In-tree build of librmtb200.so (host C++ only; device code is compiled at
run time by NVRTC for sm_100a, and offline by `nvcc` in the build check)."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "librmtb200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_library(force=False, verbose=False):
    src = os.path.join(CSRC, "rmt_capi.cpp")
    hdr = os.path.join(ROOT, "include", "rmt_b200.h")
    if not force and _newer(LIB, [src, hdr]):
        return LIB
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        raise RuntimeError("no C++ compiler found to build librmtb200.so")
    lib64 = os.path.join(CUDA_HOME, "lib64")
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-I", os.path.join(CUDA_HOME, "include"),
           src, "-o", LIB, "-L", lib64, "-lnvrtc", "-ldl", "-lpthread", "-Wl,-rpath," + lib64]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


def nvcc_cubin(model_header_text, out_dir, name, block=128, ptxas_verbose=True):
    """Offline cross-compilation of the same translation unit NVRTC sees
    (build check + `-Xptxas -v` resource report + cuobjdump)."""
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "rmt_model.cuh"), "w") as f:
        f.write(model_header_text)
    cubin = os.path.join(out_dir, name + ".cubin")
    cmd = ["nvcc"] + NVCC_ARCH + ["-lineinfo", "-O3", "-std=c++17", "-DRMT_BLOCK=%d" % block, "-I", out_dir, "-cubin",
                                  "-o", cubin, os.path.join(CSRC, "rmt_kernels.cu")]
    if ptxas_verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, check=True, capture_output=True, text=True)
    with open(os.path.join(out_dir, name + ".ptxas.txt"), "w") as f:
        f.write(res.stderr)
    return cubin, res.stderr


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
