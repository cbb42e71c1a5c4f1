"""The C-ABI boundary without a GPU: both shared libraries load, export every symbol their headers
declare, keep the wire layouts, and the compute entry points fail loudly when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import has_gpu
from pathtracer_ocl_b200 import scene as S, trace as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_libptcuda_exports_every_declared_symbol():
    names = declared("ptcuda.h", "ptc_")
    assert set(T.EXPORTS) == set(names)
    L = T.lib()
    for n in names:
        assert hasattr(L, n), f"libptcuda.so lacks {n}"
    assert b"sm_100a" in L.ptc_version()


def test_libptscene_exports_every_declared_symbol():
    L = S.lib()
    for n in declared("ptscene.h", "pts_"):
        assert hasattr(L, n), f"libptscene.so lacks {n}"


def test_wire_record_sizes_and_offsets():
    # ocltracer.go:25-96 -- sizes are also static_asserted in include/ptwire.h
    assert S.OBJECT_DTYPE.itemsize == 1024 and S.GROUP_DTYPE.itemsize == 256
    assert S.TRIANGLE_DTYPE.itemsize == 512 and S.CAMERA_DTYPE.itemsize == 256
    assert S.OBJECT_DTYPE.fields["children"][1] == 588 and S.OBJECT_DTYPE.fields["is_textured"][1] == 844
    assert S.OBJECT_DTYPE.fields["label"][1] == 849 and S.CAMERA_DTYPE.fields["inverse"][1] == 56
    header = open(os.path.join(ROOT, "include", "ptwire.h")).read()
    for size in ("1024", "256", "512"):
        assert f"== {size}" in header


def test_job_struct_layout_matches_header():
    # the ctypes mirror must have the size the C compiler gives the struct (x86-64 natural alignment)
    src = '#include "ptcuda.h"\n#include <stdio.h>\nint main(){printf("%zu %zu", sizeof(ptc_job), sizeof(ptc_stats));return 0;}\n'
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        a, b = (int(v) for v in subprocess.check_output([os.path.join(d, "s")]).split())
    assert C.sizeof(T.PtcJob) == a and C.sizeof(T.PtcStats) == b


def test_plan_rows_interleaved_tiles():
    rows = [T.plan_rows(22, r, 3, 4) for r in range(3)]
    assert list(rows[0]) == [0, 1, 2, 3, 12, 13, 14, 15] and list(rows[1]) == [4, 5, 6, 7, 16, 17, 18, 19]
    assert list(rows[2]) == [8, 9, 10, 11, 20, 21]
    assert sorted(np.concatenate(rows)) == list(range(22))
    assert list(T.plan_rows(5)) == [0, 1, 2, 3, 4]
    assert len(T.plan_rows(0)) == 0
    with pytest.raises(ValueError):
        T.plan_rows(10, 3, 3)


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_compute_calls_fail_loudly_without_a_device():
    assert T.lib().ptc_device_count() == 0 and T.list_devices() == []
    sc = S.build_scene("default", 16, 12)
    with pytest.raises(T.PtcError, match="no usable CUDA device"):
        T.render_scene(sc, 1, S.make_seeds(1, 16 * 12))
    with pytest.raises(T.PtcError, match="no usable CUDA device"):
        T.debug_noise3d(np.zeros((4, 3), np.float32))


def test_job_validation_happens_before_any_device_work():
    sc = S.build_scene("default", 16, 12)
    seeds = S.make_seeds(1, 16 * 12)
    with pytest.raises(T.PtcError, match="at most 16"):
        T.Trace(np.tile(sc.objects, 2), None, None, 0, 1, sc.camera, seeds=seeds)
    with pytest.raises(T.PtcError, match="samples"):
        T.render_scene(sc, 0, seeds)
    with pytest.raises(T.PtcError, match="precision"):
        T.render_scene(sc, 1, seeds, precision=7)
    with pytest.raises(ValueError, match="one per pixel"):
        T.render_scene(sc, 1, seeds[:-1])
    with pytest.raises(ValueError, match="1024-byte"):
        T.Trace(sc.objects[:-1], None, None, 0, 1, sc.camera, seeds=seeds)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or open it."""
    pkg = os.path.join(ROOT, "pathtracer_ocl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, f
    import subprocess
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libptcuda.so")], text=True)
    assert "oracle" not in out
