"""In-tree build of the native parts (explicit nvcc / g++ commands, no JIT cache).

  libqandle_b200.so        CUDA kernels + planner + C ABI (include/qandle_b200.h); no torch dependency
  libqandle_b200_torch.so  thin C++ torch.library layer (TORCH_LIBRARY qandle_b200::*) on top of the C ABI

Both are written next to this package (qandle_b200/) so they travel with the repo snapshot.
Usage:  python -m qandle_b200.csrc.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB_CORE = os.path.join(PKG, "libqandle_b200.so")
LIB_TORCH = os.path.join(PKG, "libqandle_b200_torch.so")

CORE_SRCS = ["capi.cu", "plan.cpp"]
CORE_DEPS = CORE_SRCS + ["kernels.cuh", "packed64.cuh", "flat64.cuh", "flat128.cuh", "exchange.cuh", "plan.h", os.path.join(ROOT, "include", "qandle_b200.h")]
TORCH_SRCS = ["torch_ops.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall", "-shared",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build the sm_100a kernels")
    return cand


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    for d in deps:
        p = d if os.path.isabs(d) else os.path.join(HERE, d)
        if os.path.getmtime(p) > t:
            return True
    return False


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"build failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout, r.stderr)


def build_core(force=False, verbose=False):
    if force or _stale(LIB_CORE, CORE_DEPS):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_CORE] + CORE_SRCS
        _run(cmd, verbose)
    return LIB_CORE


def build_torch(force=False, verbose=False):
    build_core(force, verbose)
    deps = TORCH_SRCS + [LIB_CORE, os.path.join(ROOT, "include", "qandle_b200.h")]
    if force or _stale(LIB_TORCH, deps):
        import torch
        from torch.utils import cpp_extension as ce

        inc = []
        for p in ce.include_paths() + [os.path.join(ce.CUDA_HOME or "/usr/local/cuda", "include")]:
            inc += ["-I", p]
        torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
        cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
        cmd = [
            "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}",
            "-DTORCH_API_INCLUDE_EXTENSION_H",
        ] + inc + ["-o", LIB_TORCH] + TORCH_SRCS + [
            "-L", PKG, "-lqandle_b200", "-L", torch_lib, "-lc10", "-ltorch_cpu", "-ltorch", "-lc10_cuda",
            "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}",
        ]
        _run(cmd, verbose)
    return LIB_TORCH


def build_all(force=False, verbose=False):
    build_core(force, verbose)
    build_torch(force, verbose)
    return LIB_CORE, LIB_TORCH


def build_variant(name, defines, verbose=False):
    """A second build of both libraries with extra -D macros into qandle_b200/_variants/<name>/ (load it with QB_LIB_DIR)."""
    global LIB_CORE, LIB_TORCH, PKG
    saved = (LIB_CORE, LIB_TORCH, PKG, list(NVCC_FLAGS))
    out = os.path.join(saved[2], "_variants", name)
    os.makedirs(out, exist_ok=True)
    try:
        PKG = out
        LIB_CORE = os.path.join(out, "libqandle_b200.so")
        LIB_TORCH = os.path.join(out, "libqandle_b200_torch.so")
        NVCC_FLAGS.extend(f"-D{d}" for d in defines)
        build_core(True, verbose)
        build_torch(False, verbose)
        return out
    finally:
        LIB_CORE, LIB_TORCH, PKG = saved[:3]
        NVCC_FLAGS[:] = saved[3]


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print("built variant", build_variant(sys.argv[i + 1], [a[2:] for a in sys.argv[i + 2:] if a.startswith("-D")], "--verbose" in sys.argv))
    else:
        build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
        print("built", LIB_CORE, LIB_TORCH)
