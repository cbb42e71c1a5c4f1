"""ctypes binding of libptcuda (include/ptcuda.h) -- the Python mirror of the reference's render
entry point ``ocl.Trace`` (reference internal/ocl/ocltracer.go:100) and of ``listDevices``
(cmd/pt/main.go:98-112).

``Trace(objects, triangles, groups, device_index, samples, camera, textures, sphere_textures,
cube_textures)`` keeps the reference's argument order and meaning and returns the same thing: a flat
float64 array of W*H RGBA values.  Everything runs through the C ABI; there is no Python or CPU
implementation behind it -- a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("PTCUDA_LIB") or os.path.join(_HERE, "libptcuda.so")   # override: kernel A/B experiments

PTC_ABI_VERSION = 1
FP32, FP64 = 0, 1
RNG_PARITY, RNG_FAST = 0, 1
FEATURE_NEE, FEATURE_CYLINDER_CAPS = 1, 2          # code paths the reference ships disabled (ptcuda.h)
FRAME_F64, FRAME_F32 = 0, 1
FRAME_HANDLE_BYTES = 64


class PtcJob(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("objects", C.c_void_p), ("n_objects", C.c_int32),
        ("triangles", C.c_void_p), ("n_triangles", C.c_int32),
        ("groups", C.c_void_p), ("n_groups", C.c_int32),
        ("camera", C.c_void_p),
        ("tex", C.c_void_p * 3), ("tex_w", C.c_int32 * 3), ("tex_h", C.c_int32 * 3), ("tex_layers", C.c_int32 * 3),
        ("seeds", C.c_void_p),
        ("samples", C.c_int32), ("precision", C.c_int32), ("rng_mode", C.c_int32),
        ("devices", C.POINTER(C.c_int32)), ("n_devices", C.c_int32),
        ("shard_index", C.c_int32), ("shard_count", C.c_int32), ("rows_per_tile", C.c_int32),
        ("features", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class PtcStats(C.Structure):
    _fields_ = [
        ("upload_ms", C.c_double), ("kernel_ms", C.c_double), ("read_ms", C.c_double),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("p2p_bytes", C.c_int64), ("paths", C.c_int64),
        ("kernel_launches", C.c_int32), ("n_devices", C.c_int32), ("rows", C.c_int32), ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


EXPORTS = ["ptc_device_count", "ptc_device_name", "ptc_render", "ptc_open", "ptc_trace", "ptc_read", "ptc_get_stats",
           "ptc_close", "ptc_set_seeds", "ptc_device_framebuffer", "ptc_shard_rows", "ptc_plan_rows", "ptc_version", "ptc_render_flat", "ptc_render_flat2", "ptc_trim",
           "ptc_trace_range", "ptc_reset", "ptc_read_rgba8", "ptc_read_f32", "ptc_frame_create", "ptc_frame_export",
           "ptc_frame_import", "ptc_set_frame", "ptc_frame_read", "ptc_frame_device_pointer", "ptc_frame_destroy"]

_lib = None


def lib() -> C.CDLL:
    """Load libptcuda.so.  Raises if it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "pathtracer_ocl_b200 has no CPU fallback")
        L = C.CDLL(_LIB_PATH)
        L.ptc_version.restype = C.c_char_p
        L.ptc_device_name.argtypes = [C.c_int, C.c_char_p, C.c_int]
        L.ptc_render.argtypes = [C.POINTER(PtcJob), C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_open.argtypes = [C.POINTER(PtcJob), C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        L.ptc_trace.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_read.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_get_stats.argtypes = [C.c_void_p, C.POINTER(PtcStats)]
        L.ptc_close.argtypes = [C.c_void_p]
        L.ptc_set_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_device_framebuffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int32)]
        L.ptc_shard_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int]
        L.ptc_debug_noise3d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_plan_rows.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int]
        L.ptc_trace_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int]
        L.ptc_reset.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_read_rgba8.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_read_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_frame_create.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        L.ptc_frame_export.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_frame_import.argtypes = [C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        L.ptc_set_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_frame_read.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ptc_frame_device_pointer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.ptc_frame_destroy.argtypes = [C.c_void_p]
        L.ptc_debug_dfma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_char_p, C.c_int]
        L.ptc_debug_launch_plan.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.ptc_debug_mesh_index.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_char_p, C.c_int]
        L.ptc_debug_mesh_index.restype = C.c_int64
        L.ptc_debug_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_char_p, C.c_int]
        _lib = L
    return _lib


def plan_rows(height: int, shard_index: int = 0, shard_count: int = 1, rows_per_tile: int = 0) -> np.ndarray:
    """Frame rows owned by one shard (interleaved scanline tiles); host-only, no device needed."""
    L = lib()
    n = L.ptc_plan_rows(height, rows_per_tile, shard_index, shard_count, None, 0)
    if n < 0:
        raise ValueError("bad shard arguments")
    buf = (C.c_int32 * max(n, 1))()
    L.ptc_plan_rows(height, rows_per_tile, shard_index, shard_count, buf, n)
    return np.array(buf[:n], dtype=np.int32)


class PtcError(RuntimeError):
    """The reference aborts with logrus.Fatalf on every OpenCL error; here it is an exception."""


def list_devices() -> List[str]:
    """cmd/pt/main.go:98-112: one 'Index: i Type: GPU Name: ...' line per device."""
    L = lib()
    out = []
    for i in range(L.ptc_device_count()):
        buf = C.create_string_buffer(256)
        if L.ptc_device_name(i, buf, 256) == 0:
            out.append(f"Index: {i} Type: GPU Name: {buf.value.decode()}")
    return out


def _as_bytes(a, record: int, what: str) -> np.ndarray:
    arr = np.ascontiguousarray(a)
    arr = arr.view(np.uint8).reshape(-1)
    if arr.size % record:
        raise ValueError(f"{what}: {arr.size} bytes is not a multiple of the {record}-byte record")
    return arr


def _pack_textures(images) -> Optional[np.ndarray]:
    """prepareTextures (ocltracer.go:228-254): same-sized RGBA8 images packed back to back."""
    if images is None:
        return None
    if isinstance(images, np.ndarray):
        arr = images
    else:
        if len(images) == 0:
            return None
        arr = np.stack([np.asarray(im) for im in images])
    if arr.ndim != 4 or arr.shape[-1] != 4 or arr.dtype != np.uint8:
        raise ValueError("textures must be uint8 arrays shaped [layers, height, width, 4] (RGBA)")
    return np.ascontiguousarray(arr)


class _Job:
    """Keeps the numpy buffers alive next to the ctypes struct that points into them."""

    def __init__(self, objects, triangles, groups, camera, textures, sphere_textures, cube_textures, seeds, samples,
                 precision, rng_mode, devices, shard_index, shard_count, rows_per_tile, features=0):
        self.objects = _as_bytes(objects, 1024, "objects")
        self.triangles = _as_bytes(triangles, 512, "triangles") if triangles is not None else np.zeros(0, np.uint8)
        self.groups = _as_bytes(groups, 256, "groups") if groups is not None else np.zeros(0, np.uint8)
        self.camera = _as_bytes(camera, 256, "camera")
        self.tex = [_pack_textures(t) for t in (textures, sphere_textures, cube_textures)]
        self.seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1)
        cam = np.frombuffer(self.camera.tobytes()[:8], dtype="<i4")
        self.width, self.height = int(cam[0]), int(cam[1])
        if self.seeds.size != self.width * self.height:
            raise ValueError(f"seeds: need one per pixel ({self.width * self.height}), got {self.seeds.size}")
        j = PtcJob()
        j.abi_version = PTC_ABI_VERSION
        j.objects, j.n_objects = self.objects.ctypes.data, self.objects.size // 1024
        j.triangles = self.triangles.ctypes.data if self.triangles.size else None
        j.n_triangles = self.triangles.size // 512
        j.groups = self.groups.ctypes.data if self.groups.size else None
        j.n_groups = self.groups.size // 256
        j.camera = self.camera.ctypes.data
        for k, t in enumerate(self.tex):
            if t is not None:
                j.tex[k] = t.ctypes.data
                j.tex_layers[k], j.tex_h[k], j.tex_w[k] = t.shape[0], t.shape[1], t.shape[2]
        j.seeds = self.seeds.ctypes.data
        j.samples, j.precision, j.rng_mode = int(samples), int(precision), int(rng_mode)
        if devices:
            self._dev = (C.c_int32 * len(devices))(*devices)
            j.devices, j.n_devices = self._dev, len(devices)
        j.shard_index, j.shard_count, j.rows_per_tile = int(shard_index), int(shard_count), int(rows_per_tile)
        j.features = int(features)
        self.struct = j


class Context:
    """Phase API: scene resident on the device(s); trace() and read() can be timed separately."""

    def __init__(self, objects, triangles, groups, camera, textures=None, sphere_textures=None, cube_textures=None, *,
                 seeds, samples: int = 1, precision: int = FP32, rng_mode: int = RNG_PARITY,
                 devices: Optional[Sequence[int]] = None, shard_index: int = 0, shard_count: int = 1,
                 rows_per_tile: int = 0, features: int = 0):
        self._job = _Job(objects, triangles, groups, camera, textures, sphere_textures, cube_textures, seeds, samples,
                         precision, rng_mode, list(devices) if devices else None, shard_index, shard_count, rows_per_tile,
                         features)
        self.width, self.height = self._job.width, self._job.height
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        if lib().ptc_open(C.byref(self._job.struct), C.byref(self._h), err, 512) != 0:
            raise PtcError(err.value.decode())
        n = lib().ptc_shard_rows(self._h, None, 0)
        rows = (C.c_int32 * max(n, 1))()
        lib().ptc_shard_rows(self._h, rows, n)
        self.rows = np.array(rows[:n], dtype=np.int32)

    def trace(self) -> None:
        err = C.create_string_buffer(512)
        if lib().ptc_trace(self._h, err, 512) != 0:
            raise PtcError(err.value.decode())

    def read(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = len(self.rows) * self.width * 4
        if out is None:
            out = np.empty(n, dtype=np.float64)
        if out.dtype != np.float64 or out.size != n or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a contiguous float64 array of rows*width*4 values")
        err = C.create_string_buffer(512)
        if lib().ptc_read(self._h, out.ctypes.data, err, 512) != 0:
            raise PtcError(err.value.decode())
        return out

    def read_into_ptr(self, ptr: int) -> None:
        err = C.create_string_buffer(512)
        if lib().ptc_read(self._h, C.c_void_p(ptr), err, 512) != 0:
            raise PtcError(err.value.decode())

    def set_seeds(self, seeds: np.ndarray) -> None:
        s = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1)
        if s.size != self.width * self.height:
            raise ValueError("seeds: need one per pixel")
        err = C.create_string_buffer(512)
        if lib().ptc_set_seeds(self._h, s.ctypes.data, err, 512) != 0:
            raise PtcError(err.value.decode())

    def trace_range(self, sample_begin: int, sample_end: int) -> None:
        """Progressive rendering: add samples [begin, end) of every pixel to the accumulator."""
        err = C.create_string_buffer(512)
        if lib().ptc_trace_range(self._h, sample_begin, sample_end, err, 512) != 0:
            raise PtcError(err.value.decode())

    def reset(self) -> None:
        err = C.create_string_buffer(512)
        if lib().ptc_reset(self._h, err, 512) != 0:
            raise PtcError(err.value.decode())

    def read_rgba8(self) -> np.ndarray:
        """The frame as 8-bit RGBA (clamp(round(c*255)), alpha 255), tone-mapped on the device."""
        out = np.empty((len(self.rows), self.width, 4), dtype=np.uint8)
        err = C.create_string_buffer(512)
        if lib().ptc_read_rgba8(self._h, out.ctypes.data, err, 512) != 0:
            raise PtcError(err.value.decode())
        return out

    def read_f32(self) -> np.ndarray:
        """The frame as float32 RGBA (converted on the device; half of read()'s bytes)."""
        out = np.empty((len(self.rows), self.width, 4), dtype=np.float32)
        err = C.create_string_buffer(512)
        if lib().ptc_read_f32(self._h, out.ctypes.data, err, 512) != 0:
            raise PtcError(err.value.decode())
        return out

    def set_frame(self, frame: "Optional[Frame]") -> None:
        """Attach a Frame: trace() then stores every pixel straight into it (by frame row); None detaches."""
        err = C.create_string_buffer(512)
        if lib().ptc_set_frame(self._h, frame._h if frame is not None else None, err, 512) != 0:
            raise PtcError(err.value.decode())
        self._frame = frame

    def set_seeds_ptr(self, ptr: int) -> None:
        err = C.create_string_buffer(512)
        if lib().ptc_set_seeds(self._h, C.c_void_p(ptr), err, 512) != 0:
            raise PtcError(err.value.decode())

    def stats(self) -> dict:
        st = PtcStats()
        lib().ptc_get_stats(self._h, C.byref(st))
        return st.as_dict()

    def device_framebuffer(self, local_index: int = 0):
        """(device pointer, number of doubles, CUDA device ordinal) of a local device's packed rows."""
        p, n, d = C.c_void_p(), C.c_int64(), C.c_int32()
        if lib().ptc_device_framebuffer(self._h, local_index, C.byref(p), C.byref(n), C.byref(d)) != 0:
            raise PtcError("bad local device index")
        return p.value, n.value, d.value

    def close(self) -> None:
        if self._h:
            lib().ptc_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Frame:
    """A whole-frame buffer on one device that sharded contexts render into directly (ptc_frame_*): the gather of
    SURVEY.md 8e fused into the trace kernel's epilogue as peer stores.  `Frame(device, w, h)` allocates;
    `export()` gives the 64-byte handle another process turns into its own mapping with `Frame.attach(...)`."""

    def __init__(self, device: int, width: int, height: int, fmt: int = FRAME_F64, _handle: Optional[bytes] = None):
        self.width, self.height, self.format, self.device = width, height, fmt, device
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        if _handle is None:
            rc = lib().ptc_frame_create(device, width, height, fmt, C.byref(self._h), err, 512)
        else:
            buf = C.create_string_buffer(_handle, FRAME_HANDLE_BYTES)
            rc = lib().ptc_frame_import(device, buf, width, height, fmt, C.byref(self._h), err, 512)
        if rc != 0:
            raise PtcError(err.value.decode())

    @classmethod
    def attach(cls, device: int, handle: bytes, width: int, height: int, fmt: int = FRAME_F64) -> "Frame":
        return cls(device, width, height, fmt, _handle=handle)

    def export(self) -> bytes:
        buf = C.create_string_buffer(FRAME_HANDLE_BYTES)
        err = C.create_string_buffer(512)
        if lib().ptc_frame_export(self._h, buf, err, 512) != 0:
            raise PtcError(err.value.decode())
        return buf.raw

    def read(self, out_ptr: Optional[int] = None) -> Optional[np.ndarray]:
        """[H, W, 4] float64 (float32 for FRAME_F32); with out_ptr the bytes go to that host address instead."""
        out = None
        if out_ptr is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32 if self.format == FRAME_F32 else np.float64)
            out_ptr = out.ctypes.data
        err = C.create_string_buffer(512)
        if lib().ptc_frame_read(self._h, C.c_void_p(out_ptr), err, 512) != 0:
            raise PtcError(err.value.decode())
        return out

    def device_pointer(self):
        p, n, d = C.c_void_p(), C.c_int64(), C.c_int32()
        lib().ptc_frame_device_pointer(self._h, C.byref(p), C.byref(n), C.byref(d))
        return p.value, n.value, d.value

    def close(self) -> None:
        if self._h:
            lib().ptc_frame_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def Trace(objects, triangles, groups, device_index: int, samples: int, camera, textures=None, sphere_textures=None,
          cube_textures=None, *, seeds: Optional[np.ndarray] = None, precision: int = FP32, rng_mode: int = RNG_PARITY,
          devices: Optional[Sequence[int]] = None, features: int = 0) -> np.ndarray:
    """Drop-in for ``ocl.Trace`` (ocltracer.go:100).  Positional arguments as in the reference.

    ``seeds``: one float64 in [0,1) per pixel; when omitted they are drawn like the reference does
    (``rand.Float64()`` per pixel, ocltracer.go:260-263) from numpy's global generator.
    Returns the flat ``[]float64`` of W*H*4 RGBA values."""
    cam = _as_bytes(camera, 256, "camera")
    w, h = (int(v) for v in np.frombuffer(cam.tobytes()[:8], dtype="<i4"))
    if seeds is None:
        seeds = np.random.random_sample(w * h)
    job = _Job(objects, triangles, groups, cam, textures, sphere_textures, cube_textures, seeds, samples, precision,
               rng_mode, list(devices) if devices else [int(device_index)], 0, 1, 0, features)
    out = np.empty(w * h * 4, dtype=np.float64)
    err = C.create_string_buffer(512)
    if lib().ptc_render(C.byref(job.struct), out.ctypes.data, err, 512) != 0:
        raise PtcError(err.value.decode())
    return out


def render_scene(scene, samples: int, seeds: np.ndarray, **kw) -> np.ndarray:
    """Convenience: Trace() on a pathtracer_ocl_b200.scene.SceneBuffers; returns [H, W, 4]."""
    out = Trace(scene.objects, scene.triangles if scene.n_triangles else None, scene.groups if scene.n_groups else None,
                kw.pop("device_index", 0), samples, scene.camera, scene.textures[0], scene.textures[1], scene.textures[2],
                seeds=seeds, **kw)
    return out.reshape(scene.height, scene.width, 4)


def open_scene(scene, samples: int, seeds: np.ndarray, **kw) -> Context:
    return Context(scene.objects, scene.triangles if scene.n_triangles else None, scene.groups if scene.n_groups else None,
                   scene.camera, scene.textures[0], scene.textures[1], scene.textures[2], seeds=seeds, samples=samples, **kw)


def debug_noise3d(xyz: np.ndarray, rng_mode: int = RNG_PARITY) -> np.ndarray:
    """Test hook: noise3D evaluated on the device."""
    a = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.empty(a.shape[0], dtype=np.float32)
    err = C.create_string_buffer(512)
    if lib().ptc_debug_noise3d(a.ctypes.data, a.shape[0], rng_mode, out.ctypes.data, err, 512) != 0:
        raise PtcError(err.value.decode())
    return out


def debug_fma_peak(device: int = 0) -> float:
    """Measurement hook: achieved FP32 FFMA throughput of `device` in TFLOP/s (a pure-FFMA kernel, best of 5)."""
    out = C.c_double(0.0)
    err = C.create_string_buffer(512)
    if lib().ptc_debug_fma_peak(int(device), C.byref(out), err, 512) != 0:
        raise PtcError(err.value.decode())
    return float(out.value)


def debug_dfma_peak(device: int = 0) -> float:
    """Measurement hook: achieved FP64 DFMA throughput of `device` in TFLOP/s (the fp64 mode's roofline)."""
    out = C.c_double(0.0)
    err = C.create_string_buffer(512)
    if lib().ptc_debug_dfma_peak(int(device), C.byref(out), err, 512) != 0:
        raise PtcError(err.value.decode())
    return float(out.value)


_MESH_INDEX_ARRAYS = {  # name -> (selector, dtype, trailing shape)
    "wide": (0, np.float64, (4,)), "tri_info": (4, np.int32, (2,)), "tri_test": (5, np.float64, (3, 4)),
    "node_lo": (6, np.float64, (4,)), "node_hi": (7, np.float64, (4,)), "node_parent": (8, np.int32, ()),
    "mesh": (9, np.float64, (8,)), "node_range": (10, np.int32, (2,)),
}


def debug_launch_plan(scene, samples: int, sm_count: int = 148, shard_index: int = 0, shard_count: int = 1, rows_per_tile: int = 0):
    """Test hook (host only): (slices per pixel, initial launch order of the 8x4-pixel tiles) ptc_open would plan."""
    seeds = np.zeros(scene.width * scene.height)
    job = _Job(scene.objects, scene.triangles if scene.n_triangles else None, scene.groups if scene.n_groups else None,
               scene.camera, None, None, None, seeds, samples, FP32, RNG_PARITY, None, shard_index, shard_count, rows_per_tile)
    err = C.create_string_buffer(512)
    slices = C.c_int32(0)
    n = lib().ptc_debug_launch_plan(C.byref(job.struct), sm_count, C.byref(slices), None, 0, err, 512)
    if n < 0:
        raise PtcError(err.value.decode())
    order = np.zeros(max(n, 1), dtype=np.int32)
    lib().ptc_debug_launch_plan(C.byref(job.struct), sm_count, C.byref(slices), order.ctypes.data, n, err, 512)
    return int(slices.value), order[:n]


def debug_mesh_index(scene) -> dict:
    """Test hook (host only, no device needed): the rebuilt mesh BVH and the reference-node tables the C
    layer derives from a scene's triangles/groups, in double, as numpy arrays."""
    seeds = np.zeros(scene.width * scene.height)
    job = _Job(scene.objects, scene.triangles if scene.n_triangles else None, scene.groups if scene.n_groups else None,
               scene.camera, None, None, None, seeds, 1, FP64, RNG_PARITY, None, 0, 1, 0)
    out = {}
    for name, (sel, dtype, tail) in _MESH_INDEX_ARRAYS.items():
        err = C.create_string_buffer(512)
        n = lib().ptc_debug_mesh_index(C.byref(job.struct), sel, None, 0, err, 512)
        if n < 0:
            raise PtcError(err.value.decode())
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        if n > 0 and lib().ptc_debug_mesh_index(C.byref(job.struct), sel, buf.ctypes.data, n, err, 512) < 0:
            raise PtcError(err.value.decode())
        out[name] = buf[:n].view(dtype).reshape((-1,) + tail)
    return out
