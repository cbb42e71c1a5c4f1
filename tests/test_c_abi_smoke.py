"""tests/c_abi_smoke.c: a plain C program that calls ptc_render_flat2 with exactly the argument shapes the cgo binding
(go/internal/cuda/cuda.go) passes -- the stand-in for compiling the Go binding in an image without a Go toolchain.
Without a GPU it must be refused loudly; on the B200 box it must render a plausible picture."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pathtracer_ocl_b200")


def build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-L", PKG, "-lptcuda", "-lm", f"-Wl,-rpath,{PKG}"])
    return exe


def run(exe):
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    fields = dict(kv.split("=", 1) for kv in res.stdout.strip().split(" ", 4))
    return res.returncode, fields


def test_c_caller_compiles_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    from conftest import has_gpu
    code, f = run(build(tmp_path))
    assert code == 0, f
    if not has_gpu():
        assert f["rc"] == "1" and f["devices"] == "0" and "no usable CUDA device" in f["msg"] and "no CPU fallback" in f["msg"]


@pytest.mark.gpu
def test_c_caller_renders_through_ptc_render_flat2(tmp_path):
    code, f = run(build(tmp_path))
    assert code == 0 and f["rc"] == "0" and int(f["devices"]) >= 1, f
    assert float(f["center"]) > 0.05 and float(f["corner"]) > 0.01, f


def test_go_binding_does_not_import_the_opencl_package():
    """INTEGRATION.md promises that internal/cuda builds without any OpenCL header: the binding must not import
    internal/ocl (whose ocltracer.go imports go-opencl/cl) and must only call entry points the header declares."""
    import re
    src = open(os.path.join(ROOT, "go", "internal", "cuda", "cuda.go")).read()
    imports = re.search(r"import \((.*?)\)", src, re.S).group(1)
    assert "internal/ocl" not in imports and "go-opencl" not in imports
    header = open(os.path.join(ROOT, "include", "ptcuda.h")).read()
    for sym in set(re.findall(r"C\.(ptc_\w+)\(", src)):
        assert re.search(r"\b%s\s*\(" % sym, header), sym
    for const in set(re.findall(r"C\.(PTC_\w+)", src)):
        assert re.search(r"#define\s+%s\b" % const, header), const
