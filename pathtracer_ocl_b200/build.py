"""In-tree build recipes: libptcuda.so (nvcc, sm_100a), libptscene.so and the `pt` CLI (g++).

Artifacts land next to this file so they travel with a repo snapshot.  Rebuilds only what is stale.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-Xptxas", "-v"]
CXX_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log=None):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError(f"build step failed: {' '.join(cmd)}")
    return res.stdout


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_cuda(force: bool = False) -> str:
    target = os.path.join(_HERE, "libptcuda.so")
    deps = [os.path.join(CSRC, "ptcuda.cu")] + glob.glob(os.path.join(CSRC, "kernels", "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _stale(target, deps):
        os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
        _run([nvcc_path()] + NVCC_FLAGS + ["-o", target, os.path.join(CSRC, "ptcuda.cu")],
             log=os.path.join(ROOT, "build", "ptcuda_ptxas.log"))
    return target


def build_scene_lib(force: bool = False) -> str:
    target = os.path.join(_HERE, "libptscene.so")
    host = os.path.join(CSRC, "host")
    srcs = [os.path.join(host, f) for f in ("shapes.cpp", "scenes.cpp", "ptscene_capi.cpp")]
    deps = srcs + glob.glob(os.path.join(host, "*.hpp")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _stale(target, deps):
        _run(["g++"] + CXX_FLAGS + ["-shared", "-o", target] + srcs)
    return target


def build_cli(force: bool = False) -> str:
    """`pt`: C++ stand-in for the reference's cmd/pt (same flags), linked against both libraries."""
    target = os.path.join(_HERE, "pt")
    src = os.path.join(CSRC, "host", "pt_main.cpp")
    if not os.path.exists(src):
        return ""
    deps = [src, os.path.join(_HERE, "libptcuda.so"), os.path.join(_HERE, "libptscene.so")]
    if force or _stale(target, deps):
        _run(["g++"] + CXX_FLAGS + ["-o", target, src, "-L" + _HERE, "-lptscene", "-lptcuda", "-Wl,-rpath,$ORIGIN"])
    return target


def build_all(force: bool = False) -> None:
    build_scene_lib(force)
    build_cuda(force)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", ", ".join(sorted(os.path.basename(p) for p in glob.glob(os.path.join(_HERE, "*.so")))))
