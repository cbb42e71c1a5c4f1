"""ctypes binding of the CPU oracle (oracle/tracer_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tracer_oracle.cpp")
    deps = [src, os.path.join(_HERE, "canon_rng.h"), os.path.join(_HERE, "..", "include", "ptwire.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(_LIB_PATH) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_counter_name.restype = C.c_char_p
        L.oracle_trace3.restype = C.c_int
        L.oracle_trace3.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]
        L.oracle_noise3d_array_mode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_noise3d.restype = C.c_float
        L.oracle_noise3d.argtypes = [C.c_float] * 3
        L.oracle_sinf.restype = C.c_float
        L.oracle_sinf.argtypes = [C.c_float]
        L.oracle_noise3d_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_sinf_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_ray_box.argtypes = [C.c_void_p] * 4
        L.oracle_spherical_map.argtypes = [C.c_void_p] * 2
        L.oracle_cube_uv.argtypes = [C.c_void_p] * 2
        L.oracle_sunflower.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.oracle_read_imagef.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_void_p]
        L.oracle_schlick.restype = C.c_double
        L.oracle_schlick.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        _lib = L
    return _lib


def counter_names():
    L = lib()
    return [L.oracle_counter_name(i).decode() for i in range(L.oracle_counter_count())]


NEE, CYLINDER_CAPS = 1, 2      # `features` bits: the code paths the reference ships commented out (tracer.cl:1168, :437-444)


def trace(scene, seeds: np.ndarray, samples: int, precision: int = 1, rows: Optional[Tuple[int, int]] = None,
          nthreads: Optional[int] = None, rng_mode: int = 0, features: int = 0) -> Tuple[np.ndarray, Dict[str, int]]:
    """Render rows [rows[0], rows[1]) of `scene` (a pathtracer_ocl_b200.scene.SceneBuffers).

    Returns (rgba float64 [nrows, W, 4], event counters)."""
    L = lib()
    W, H = scene.width, scene.height
    seeds = np.ascontiguousarray(seeds, dtype=np.float64)
    assert seeds.size == W * H, "one seed per pixel"
    r0, r1 = rows if rows is not None else (0, H)
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    out = np.zeros(((r1 - r0), W, 4), dtype=np.float64)
    ptrs = (C.c_void_p * 3)()
    tw, th, tl = (C.c_int32 * 3)(), (C.c_int32 * 3)(), (C.c_int32 * 3)()
    keep = []
    for c in range(3):
        t = scene.textures[c]
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.uint8)
            keep.append(t)
            ptrs[c] = t.ctypes.data
            tl[c], th[c], tw[c] = t.shape[0], t.shape[1], t.shape[2]
    counters = np.zeros(L.oracle_counter_count(), dtype=np.uint64)
    rc = L.oracle_trace3(scene.objects.ctypes.data, scene.n_objects,
                         scene.triangles.ctypes.data if scene.n_triangles else None, scene.n_triangles,
                         scene.groups.ctypes.data if scene.n_groups else None, scene.n_groups,
                         scene.camera.ctypes.data, ptrs, tw, th, tl, seeds.ctypes.data, samples, precision, rng_mode, features, r0, r1,
                         nthreads, out.ctypes.data, counters.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_trace failed with code {rc}")
    return out, dict(zip(counter_names(), (int(v) for v in counters)))


# Fixed per-event weights of the cost model, in algorithmic flops.
FLOP_WEIGHTS = {
    "paths": 95.0, "obj_plane": 56.0 + 4.0, "obj_sphere": 56.0 + 40.0, "obj_cylinder": 56.0 + 33.0,
    "obj_cube": 56.0 + 28.0, "obj_group": 56.0, "box_tests": 26.0,
    # triangle tests by exit stage: det 16 / u 28 / v 46 / full 52 (+17 recorded) -> incremental weights
    "tri_det": 16.0, "tri_u": 12.0, "tri_v": 18.0, "tri_full": 6.0, "tri_recorded": 17.0,
    "shaded": 90.0 + 13.0, "diffuse": 81.0, "mirror": 25.0, "refract": 55.0, "thin_pass": 25.0, "tex_lookups": 40.0,
}


def model_flops(counters: Dict[str, int], dof: bool = False) -> float:
    f = sum(FLOP_WEIGHTS[k] * counters.get(k, 0) for k in FLOP_WEIGHTS)
    if dof:
        f += 30.0 * counters.get("paths", 0)
    return f


# ---- the reference's own kernel, compiled for the CPU (oracle/build_ref.py) -----------------------------------------
_ref_libs: Dict[int, Optional[C.CDLL]] = {}


def ref_lib(features: int = 0) -> Optional[C.CDLL]:
    """oracle/_ref/libtracer_ref.so: /root/reference/internal/ocl/tracer.cl compiled as C++ through cl_shim.hpp
    (features != 0: the variant with the NEE call / the cylinder caps un-commented, libtracer_ref_f<features>.so).
    Built here when the reference source is present; elsewhere (the GPU box) the prebuilt library is used; None if
    neither exists."""
    if features not in _ref_libs:
        from . import build_ref
        path = build_ref.build(features=features)
        if not path:
            _ref_libs[features] = None
            return None
        L = C.CDLL(path)
        L.ref_trace.restype = C.c_int
        L.ref_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p),
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_void_p]
        L.ref_closest.restype = C.c_int
        L.ref_closest.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _ref_libs[features] = L
    return _ref_libs[features]


def ref_closest(scene, rays: np.ndarray) -> np.ndarray:
    """The reference's own findClosestIntersection (tracer.cl:537-742) for world-space rays [n, 6] (origin, direction).
    Returns [n, 8]: t, object index (-1 = none), recorded normal xyz, recorded colour rgb."""
    L = ref_lib()
    if L is None:
        raise RuntimeError("the compiled reference kernel is not available")
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    out = np.zeros((rays.shape[0], 8), dtype=np.float64)
    L.ref_closest(scene.objects.ctypes.data, scene.n_objects, scene.triangles.ctypes.data if scene.n_triangles else None,
                  scene.n_triangles, scene.groups.ctypes.data if scene.n_groups else None, scene.n_groups, rays.ctypes.data,
                  rays.shape[0], out.ctypes.data)
    return out


def ref_trace(scene, seeds: np.ndarray, samples: int, rows: Optional[Tuple[int, int]] = None,
              nthreads: Optional[int] = None, features: int = 0) -> np.ndarray:
    """Rows [rows[0], rows[1]) of `scene` rendered by the REFERENCE kernel itself (fp64, canonical sin).  [nrows, W, 4]."""
    L = ref_lib(features)
    if L is None:
        raise RuntimeError("the compiled reference kernel is not available (no /root/reference and no oracle/_ref/libtracer_ref.so)")
    W, H = scene.width, scene.height
    seeds = np.ascontiguousarray(seeds, dtype=np.float64)
    assert seeds.size == W * H, "one seed per pixel"
    r0, r1 = rows if rows is not None else (0, H)
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    out = np.zeros(((r1 - r0), W, 4), dtype=np.float64)
    ptrs = (C.c_void_p * 3)()
    tw, th, tl = (C.c_int32 * 3)(), (C.c_int32 * 3)(), (C.c_int32 * 3)()
    keep = []
    for c in range(3):
        t = scene.textures[c]
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.uint8)
            keep.append(t)
            ptrs[c] = t.ctypes.data
            tl[c], th[c], tw[c] = t.shape[0], t.shape[1], t.shape[2]
    rc = L.ref_trace(scene.objects.ctypes.data, scene.n_objects,
                     scene.triangles.ctypes.data if scene.n_triangles else None, scene.n_triangles,
                     scene.groups.ctypes.data if scene.n_groups else None, scene.n_groups,
                     scene.camera.ctypes.data, ptrs, tw, th, tl, seeds.ctypes.data, samples, r0, r1, nthreads, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"ref_trace failed with code {rc}")
    return out
