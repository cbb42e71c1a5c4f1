"""TEST INFRASTRUCTURE.  Compiles the reference's OWN device kernel for the host CPU.

    python oracle/build_ref.py            # -> oracle/_ref/libtracer_ref.so (git-ignored; travels with gpurun snapshots)

Source: /root/reference/internal/ocl/tracer.cl, read where it lies (nothing of it is copied into the repository).
OpenCL C is compiled as C++ through oracle/cl_shim.hpp.  One construct has no C++ spelling -- the vector literal
`(double4)(a, b, c, d)` -- so those (38 occurrences: double4, double2, float4) are rewritten to `mk_double4(a, b, c, d)`
in memory; the translation unit exists only in a temporary directory during the compile.  Nothing else is touched:
every line of geometry, shading and control flow that runs is the reference's.

The GPU box has no /root/reference: there only the prebuilt library is used (tests skip when it is absent).
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_KERNEL = "/root/reference/internal/ocl/tracer.cl"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_LIB = os.path.join(OUT_DIR, "libtracer_ref.so")
CXXFLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-pthread", "-w"]


def available() -> bool:
    return os.path.exists(OUT_LIB)


def translate(text: str) -> str:
    """OpenCL vector literals -> constructor calls; the only edit made to the reference source."""
    return re.sub(r"\((double4|double2|float4)\)\s*\(", lambda m: f"mk_{m.group(1)}(", text)


def build(force: bool = False) -> str:
    """Returns the library path, or "" when neither the reference source nor a prebuilt library is around."""
    deps = [os.path.join(HERE, f) for f in ("cl_shim.hpp", "ref_driver.cpp", "canon_rng.h", "build_ref.py")]
    if not os.path.exists(REFERENCE_KERNEL):
        return OUT_LIB if available() else ""
    deps.append(REFERENCE_KERNEL)
    if not force and available() and all(os.path.getmtime(d) <= os.path.getmtime(OUT_LIB) for d in deps):
        return OUT_LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(REFERENCE_KERNEL, "r", encoding="utf-8") as f:
        kernel = translate(f.read())
    with tempfile.TemporaryDirectory(prefix="ptref_") as tmp:
        with open(os.path.join(tmp, "tracer_cl.inc"), "w", encoding="utf-8") as f:
            f.write(kernel)
        cmd = ["g++"] + CXXFLAGS + ["-shared", "-I", HERE, "-I", tmp, "-I", os.path.join(HERE, "..", "include"),
                                    "-o", OUT_LIB, os.path.join(HERE, "ref_driver.cpp")]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout)
            raise RuntimeError("compiling the reference kernel for the CPU failed")
    return OUT_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv) or "reference source not present and no prebuilt library")
