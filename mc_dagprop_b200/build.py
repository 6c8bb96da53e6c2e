"""In-tree build of the native pieces (no JIT cache: the built .so files travel with the repo).

    libmcdp_b200.so                      C ABI (include/mcdp_b200.h): host plan compiler + sm_100a kernels
    monte_carlo/_core<EXT_SUFFIX>        pybind11 module with the reference's Python surface

Run as ``python -m mc_dagprop_b200.build`` or through ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmcdp_b200.so")
NVCC = os.environ.get("MCDP_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CXX = os.environ.get("MCDP_HOST_CXX", "/usr/bin/g++")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*exts: str) -> list[str]:
    out = [os.path.join(ROOT, "include", "mcdp_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(exts):
            out.append(os.path.join(CSRC, f))
    return out


def build_lib(force: bool = False, verbose: bool = False) -> str:
    deps = _sources(".cu", ".cuh", ".cpp", ".hpp", ".h", ".inl")
    deps = [d for d in deps if not d.endswith("pybind_core.cpp")]
    if force or _newer(LIB, deps):
        tmp = LIB + ".tmp"  # compiled beside the target and renamed: a process that has the old file mapped keeps it
        cmd = [NVCC, "-O3", "-std=c++17", *ARCH_FLAGS, "-lineinfo", "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-O3,-Wall",
               "-shared", "-cudart", "static", "-o", tmp, *os.environ.get("MCDP_NVCC_EXTRA", "").split(),
               os.path.join(CSRC, "mcdp_capi.cu"), os.path.join(CSRC, "mcdp_analytic.cu"), os.path.join(CSRC, "mcdp_plan.cpp")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
        os.replace(tmp, LIB)
    return LIB


def core_path() -> str:
    return os.path.join(PKG, "monte_carlo", "_core" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_core(force: bool = False) -> str:
    import pybind11

    src = os.path.join(CSRC, "pybind_core.cpp")
    out = core_path()
    if not os.path.exists(src):
        return out
    if force or _newer(out, [src, os.path.join(ROOT, "include", "mcdp_b200.h"), LIB]):
        cmd = [HOST_CXX, "-O2", "-std=c++20", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
               f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", f"-I{os.path.join(ROOT, 'include')}",
               "-o", out + ".tmp", src, f"-L{PKG}", "-lmcdp_b200", "-Wl,-rpath,$ORIGIN/.."]
        subprocess.run(cmd, check=True)
        os.replace(out + ".tmp", out)
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_core(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
