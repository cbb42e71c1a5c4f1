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


def available(features: int = 0) -> bool:
    return os.path.exists(lib_path(features))


def translate(text: str) -> str:
    """OpenCL vector literals -> constructor calls; the only edit made to the reference source."""
    return re.sub(r"\((double4|double2|float4)\)\s*\(", lambda m: f"mk_{m.group(1)}(", text)


# The two code paths the reference ships commented out.  A feature variant of the library is the same
# source with exactly these comment markers removed, in memory: nothing else differs from the stock build.
NEE, CYLINDER_CAPS = 1, 2
FEATURE_LINES = {
    NEE: ["            // nextEventEstimation(objects, numObjects, groups, triangles, &bnce, fgi, fgi2, n, mask, x, &accumColor);"],
    CYLINDER_CAPS: ["//    double2 caps = intersectCaps(tRayOrigin, tRayDirection, obj.minY, obj.maxY);", "//    if (caps.x > 0.0) {",
                    "//        out.z = caps.x;", "//    }", "//    if (caps.y > 0.0) {", "//        out.w = caps.y;", "//    }"],
}


def enable(text: str, features: int) -> str:
    """Un-comment the disabled call sites of the requested features (tracer.cl:1168 and :437-444)."""
    for bit, lines in FEATURE_LINES.items():
        if not features & bit:
            continue
        block = "\n".join(lines)
        if text.count(block) != 1:
            raise RuntimeError(f"feature {bit}: the commented-out block was not found exactly once in {REFERENCE_KERNEL}")
        live = "\n".join(re.sub(r"^(\s*)//\s?", r"\1", ln) for ln in lines)
        text = text.replace(block, live)
    return text


def lib_path(features: int = 0) -> str:
    return OUT_LIB if not features else os.path.join(OUT_DIR, f"libtracer_ref_f{features}.so")


def build(force: bool = False, features: int = 0) -> str:
    """Returns the library path, or "" when neither the reference source nor a prebuilt library is around.
    features: 0 = the kernel as shipped; NEE / CYLINDER_CAPS bits = the variants with those call sites un-commented."""
    out_lib = lib_path(features)
    deps = [os.path.join(HERE, f) for f in ("cl_shim.hpp", "ref_driver.cpp", "canon_rng.h", "build_ref.py")]
    if not os.path.exists(REFERENCE_KERNEL):
        return out_lib if available(features) else ""
    deps.append(REFERENCE_KERNEL)
    if not force and available(features) and all(os.path.getmtime(d) <= os.path.getmtime(out_lib) for d in deps):
        return out_lib
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(REFERENCE_KERNEL, "r", encoding="utf-8") as f:
        kernel = translate(enable(f.read(), features))
    with tempfile.TemporaryDirectory(prefix="ptref_") as tmp:
        with open(os.path.join(tmp, "tracer_cl.inc"), "w", encoding="utf-8") as f:
            f.write(kernel)
        cmd = ["g++"] + CXXFLAGS + ["-shared", "-I", HERE, "-I", tmp, "-I", os.path.join(HERE, "..", "include"),
                                    "-o", out_lib, os.path.join(HERE, "ref_driver.cpp")]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout)
            raise RuntimeError("compiling the reference kernel for the CPU failed")
    return out_lib


def build_all(force: bool = False):
    return [build(force, f) for f in (0, NEE, CYLINDER_CAPS, NEE | CYLINDER_CAPS)]


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv) or "reference source not present and no prebuilt library")
