"""Build the ttb200 CUDA library in-tree (sm_100a only).

    python -m torchtree_b200.build [--force]

Produces torchtree_b200/lib/libttb200.so with explicit nvcc invocations
(`-gencode arch=compute_100a,code=sm_100a -lineinfo`).  nvcc cross-compiles
without a GPU, so this runs in the authoring container; the built library
travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
BUILD = os.path.join(PKG, "_obj")
LIB = os.path.join(LIBDIR, "libttb200.so")
SOURCES = ["api.cu", "kernels_s4.cu", "kernels_small.cu", "kernels_gen.cu",
           "kernels_gmma.cu", "patterns.cu", "heights.cu", "eigen.cu", "kernels_gwarp.cu", "coalescent.cu",
           "expm.cu"]
HEADERS = [os.path.join(CSRC, "engine.cuh"), os.path.join(PKG, "..", "include", "ttb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",  # fp64 fma contraction only; no fast-math, no ftz
] + os.environ.get("TTB2_NVCC_EXTRA", "").split()   # e.g. -DTTB2_GM_TRACE (tools/profile_eval.py --trace)


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libttb200.so")
    return nvcc


def _digest() -> str:
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(path, "rb") as fp:
            h.update(fp.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libttb200.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as fp:
            if fp.read().strip() == digest:
                return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    # static CUDA runtime (nvcc default): no dependence on which libcudart the
    # host process (e.g. torch) happens to have loaded
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                  "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fp:
        fp.write(digest)
    return LIB


TORCH_EXT = os.path.join(PKG, "_ttb200_torch.so")
TORCH_EXT_SRC = os.path.join(CSRC, "torch_ext.cpp")


def build_torch_extension(force: bool = False) -> str:
    """Compile csrc/torch_ext.cpp (the torch C++ autograd Functions over the C ABI) into
    torchtree_b200/_ttb200_torch.so with an explicit g++ command: host code only, linked
    against lib/libttb200.so through an $ORIGIN-relative rpath."""
    import torch
    from torch.utils import cpp_extension

    build()
    h = hashlib.sha256()
    for path in (TORCH_EXT_SRC, HEADERS[1]):
        with open(path, "rb") as fp:
            h.update(fp.read())
    h.update(torch.__version__.encode())
    digest = h.hexdigest()
    stamp = os.path.join(LIBDIR, "_ttb200_torch.stamp")
    if not force and os.path.exists(TORCH_EXT) and os.path.exists(stamp):
        with open(stamp) as fp:
            if fp.read().strip() == digest:
                return TORCH_EXT
    import sysconfig

    cxx = os.environ.get("CXX") or shutil.which("g++")
    if not cxx:
        raise RuntimeError("g++ not found; cannot build the torch extension")
    cuda_home = os.environ.get("CUDA_HOME") or os.path.dirname(os.path.dirname(_nvcc()))
    incs = cpp_extension.include_paths() + [sysconfig.get_paths()["include"],
                                            os.path.join(PKG, "..", "include"),
                                            os.path.join(cuda_home, "include")]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-DTORCH_EXTENSION_NAME=_ttb200_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-isystem" + i for i in incs]
    cmd += [TORCH_EXT_SRC, "-o", TORCH_EXT,
            "-L" + LIBDIR, "-lttb200", "-Wl,-rpath,$ORIGIN/lib",
            "-L" + torch_lib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python",
            "-Wl,-rpath," + torch_lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the torch extension failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fp:
        fp.write(digest)
    return TORCH_EXT


def main():
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
    print(build_torch_extension(force="--force" in sys.argv))


if __name__ == "__main__":
    main()
