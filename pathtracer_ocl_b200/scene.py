"""ctypes binding of libptscene (include/ptscene.h): the host-side frontend that builds scenes.

Mirrors, for Python callers, what the reference's Go frontend does before it calls ``ocl.Trace``:
``scenes.<Name>()`` + ``ocl.BuildSceneBufferCL`` (reference internal/ocl/scene.go:14) + the camera
record of internal/app/tracer/renderer.go:44-56.  The buffers are returned as numpy byte arrays
with the exact wire layout of include/ptwire.h.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
ASSETS_DIR = os.path.join(REPO_ROOT, "assets")
_LIB_PATH = os.path.join(_HERE, "libptscene.so")

OBJECT_BYTES, TRIANGLE_BYTES, GROUP_BYTES, CAMERA_BYTES = 1024, 512, 256, 256

# numpy views of the wire records (include/ptwire.h) -- handy for tests and debugging
OBJECT_DTYPE = np.dtype({
    "names": ["transform", "inverse", "inverse_transpose", "color", "emission", "refractive_index", "type",
              "min_y", "max_y", "reflectivity", "texture_scale_x", "texture_scale_y", "texture_scale_x_nm",
              "texture_scale_y_nm", "bb_min", "bb_max", "child_count", "children", "is_textured", "texture_index",
              "is_textured_nm", "texture_index_nm", "is_env_map", "label"],
    "formats": [("<f8", 16), ("<f8", 16), ("<f8", 16), ("<f8", 4), ("<f8", 4), "<f8", "<i8", "<f8", "<f8", "<f8",
                "<f8", "<f8", "<f8", "<f8", ("<f8", 4), ("<f8", 4), "<i4", ("<i4", 64), "u1", "u1", "u1", "u1", "u1",
                "S8"],
    "offsets": [0, 128, 256, 384, 416, 448, 456, 464, 472, 480, 488, 496, 504, 512, 520, 552, 584, 588, 844, 845,
                846, 847, 848, 849],
    "itemsize": 1024,
})
GROUP_DTYPE = np.dtype({
    "names": ["bb_min", "bb_max", "color", "emission", "tri_offset", "tri_count", "child_group_count", "children"],
    "formats": [("<f8", 4), ("<f8", 4), ("<f8", 4), ("<f8", 4), "<i4", "<i4", "<i4", ("<i4", 2)],
    "offsets": [0, 32, 64, 96, 128, 132, 136, 140],
    "itemsize": 256,
})
TRIANGLE_DTYPE = np.dtype({
    "names": ["p1", "p2", "p3", "e1", "e2", "n1", "n2", "n3", "color"],
    "formats": [("<f8", 4)] * 9,
    "offsets": [0, 32, 64, 96, 128, 160, 192, 224, 256],
    "itemsize": 512,
})
CAMERA_DTYPE = np.dtype({
    "names": ["width", "height", "fov", "pixel_size", "half_width", "half_height", "aperture", "focal_length",
              "inverse"],
    "formats": ["<i4", "<i4", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", ("<f8", 16)],
    "offsets": [0, 4, 8, 16, 24, 32, 40, 48, 56],
    "itemsize": 256,
})

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_LIB_PATH)
        L.pts_scene_name.restype = C.c_char_p
        L.pts_scene_build.restype = C.c_void_p
        L.pts_scene_build.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_char_p, C.c_int32,
                                      C.c_char_p, C.c_int]
        L.pts_scene_from_obj.restype = C.c_void_p
        L.pts_scene_from_obj.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int]
        L.pts_scene_free.argtypes = [C.c_void_p]
        L.pts_scene_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        for fn in ("pts_scene_objects", "pts_scene_triangles", "pts_scene_groups", "pts_scene_camera"):
            getattr(L, fn).restype = C.c_void_p
            getattr(L, fn).argtypes = [C.c_void_p]
        L.pts_scene_texture.restype = C.c_int32
        L.pts_scene_texture.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32)]
        L.pts_fill_seeds.argtypes = [C.c_uint64, C.c_void_p, C.c_int64]
        L.pts_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32]
        L.pts_write_raw.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32]
        L.pts_obj_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.pts_mat_transform.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
        _lib = L
    return _lib


def scene_names() -> List[str]:
    """cmd/pt/main.go:92-96 listScenes()."""
    L = lib()
    return [L.pts_scene_name(i).decode() for i in range(L.pts_scene_count())]


@dataclass
class SceneBuffers:
    """What the Go frontend hands to ocl.Trace (ocltracer.go:100): packed records + textures."""
    name: str
    width: int
    height: int
    objects: np.ndarray        # uint8 [n_objects*1024]
    triangles: np.ndarray      # uint8 [n_triangles*512]
    groups: np.ndarray         # uint8 [n_groups*256]
    camera: np.ndarray         # uint8 [256]
    textures: List[Optional[np.ndarray]] = field(default_factory=lambda: [None, None, None])  # [layers,h,w,4] u8

    @property
    def n_objects(self) -> int:
        return self.objects.size // OBJECT_BYTES

    @property
    def n_triangles(self) -> int:
        return self.triangles.size // TRIANGLE_BYTES

    @property
    def n_groups(self) -> int:
        return self.groups.size // GROUP_BYTES

    def objects_view(self) -> np.ndarray:
        return self.objects.view(OBJECT_DTYPE)

    def groups_view(self) -> np.ndarray:
        return self.groups.view(GROUP_DTYPE)

    def triangles_view(self) -> np.ndarray:
        return self.triangles.view(TRIANGLE_DTYPE)

    def camera_view(self) -> np.ndarray:
        return self.camera.view(CAMERA_DTYPE)


def _copy(ptr, nbytes: int) -> np.ndarray:
    if not ptr or nbytes == 0:
        return np.zeros(0, dtype=np.uint8)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)).copy()


def _harvest(handle, name: str, width: int, height: int) -> SceneBuffers:
    L = lib()
    no, nt, ng = C.c_int32(), C.c_int32(), C.c_int32()
    L.pts_scene_counts(handle, C.byref(no), C.byref(nt), C.byref(ng))
    out = SceneBuffers(
        name=name, width=width, height=height,
        objects=_copy(L.pts_scene_objects(handle), no.value * OBJECT_BYTES),
        triangles=_copy(L.pts_scene_triangles(handle), nt.value * TRIANGLE_BYTES),
        groups=_copy(L.pts_scene_groups(handle), ng.value * GROUP_BYTES),
        camera=_copy(L.pts_scene_camera(handle), CAMERA_BYTES),
    )
    for cls in range(3):
        p, w, h = C.c_void_p(), C.c_int32(), C.c_int32()
        layers = L.pts_scene_texture(handle, cls, C.byref(p), C.byref(w), C.byref(h))
        if layers:
            out.textures[cls] = _copy(p.value, layers * w.value * h.value * 4).reshape(layers, h.value, w.value, 4)
    return out


def build_scene(name: str = "default", width: int = 640, height: int = 480, aperture: float = 0.0,
                focal_length: float = 0.0, assets_dir: str = ASSETS_DIR, tex_scale: int = 1) -> SceneBuffers:
    """Scene factory + BuildSceneBufferCL.  Defaults follow cmd/pt/main.go:47-52."""
    L = lib()
    err = C.create_string_buffer(512)
    h = L.pts_scene_build(name.encode(), width, height, aperture, focal_length, assets_dir.encode(), tex_scale, err, 512)
    if not h:
        raise RuntimeError(f"scene '{name}': {err.value.decode()}")
    try:
        return _harvest(h, name, width, height)
    finally:
        L.pts_scene_free(h)


def scene_from_obj(text: str, mtl_dir: str = "", vertex_normals: bool = False, divide_threshold: int = 0):
    """Parse OBJ text -> single-group scene buffers + (vertices, normals, groups, triangles, ignored lines) counts."""
    L = lib()
    err = C.create_string_buffer(512)
    h = L.pts_scene_from_obj(text.encode(), mtl_dir.encode(), int(vertex_normals), divide_threshold, err, 512)
    if not h:
        raise RuntimeError(err.value.decode())
    try:
        stats = (C.c_int32 * 5)()
        L.pts_obj_stats(h, stats)
        return _harvest(h, "obj", 4, 4), tuple(stats)
    finally:
        L.pts_scene_free(h)


def make_seeds(seed: int, n: int) -> np.ndarray:
    """splitmix64 stream -> doubles in [0,1), one per pixel."""
    out = np.empty(n, dtype=np.float64)
    lib().pts_fill_seeds(C.c_uint64(seed & (2**64 - 1)), out.ctypes.data, n)
    return out


def write_png(path: str, rgba: np.ndarray, width: int, height: int) -> None:
    a = np.ascontiguousarray(rgba, dtype=np.float64)
    if lib().pts_write_png(path.encode(), a.ctypes.data, width, height) != 0:
        raise OSError(f"cannot write {path}")


def write_raw(path: str, rgba: np.ndarray, width: int, height: int) -> None:
    a = np.ascontiguousarray(rgba, dtype=np.float64)
    if lib().pts_write_raw(path.encode(), a.ctypes.data, width, height) != 0:
        raise OSError(f"cannot write {path}")
