"""The CUDA path at BASELINE.json's exact configurations (parity holes): config 1 as specified
(640x480 at 1 spp, whole frame, against the reference's own kernel), config 4 at 1280x960, and the config-5 family at
3840x2160 with FULL-SIZE synthetic textures (4096x2048 sphere maps, 4096x3072 cube cross, 2048x2048x4 plane layers) in
both precisions -- strips of rows against the oracle, since a CPU frame at 4K would take minutes.

Gates are BASELINE.json's: >= 99.9 % of pixels within 1e-3 (fp32 mode) / 1e-6 (fp64 mode) at low spp."""
import numpy as np
import pytest

from oracle import oracle as O
from pathtracer_ocl_b200 import scene as S, trace as T

pytestmark = pytest.mark.gpu
TOL = {T.FP32: 1e-3, T.FP64: 1e-6}


def frac_within(img, ref, tol):
    err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
    return float((err <= tol).mean()), float(np.nanmax(err))


@pytest.mark.skipif(O.ref_lib() is None, reason="no prebuilt oracle/_ref/libtracer_ref.so")
def test_config1_default_640x480_at_1spp_whole_frame_against_the_reference_kernel():
    W, H = 640, 480
    sc = S.build_scene("default", W, H)
    seeds = S.make_seeds(0x5EED0001, W * H)
    ref = O.ref_trace(sc, seeds, 1)                       # the reference's tracer.cl, compiled for the CPU
    orc, _ = O.trace(sc, seeds, 1, precision=1)
    assert np.array_equal(ref, orc)                       # ... which the oracle restates bit for bit
    for precision in (T.FP64, T.FP32):
        img = T.render_scene(sc, 1, seeds, precision=precision)
        frac, worst = frac_within(img, ref, TOL[precision])
        assert frac >= 0.999, f"precision {precision}: {frac * 100:.3f}% within {TOL[precision]:g} (worst {worst:.3e})"
        assert np.all(img[..., 3] == 1.0)


@pytest.mark.parametrize("precision", [T.FP32, T.FP64], ids=["fp32", "fp64"])
def test_config4_gopher_1280x960_strip(precision):
    W, H = 1280, 960
    sc = S.build_scene("gopher", W, H)
    seeds = S.make_seeds(0x5EED0004, W * H)
    img = T.render_scene(sc, 1, seeds, precision=precision)
    for rows in ((520, 532), (700, 708)):                 # through the gopher's body / its feet and the floor
        ref, _ = O.trace(sc, seeds, 1, 1, rows=rows)
        frac, worst = frac_within(img[rows[0]:rows[1]], ref, TOL[precision])
        assert frac >= 0.999, (rows, frac, worst)


@pytest.mark.parametrize("name,rows", [("textures", (1500, 1506)), ("textures", (700, 704)), ("envmap", (300, 306)),
                                       ("envmap", (1300, 1304)), ("cubemap", (1200, 1204)), ("cubemap", (400, 404))])
def test_config5_3840x2160_full_size_textures_strips(name, rows):
    """Full-size texture arrays on the device: 64 MB of plane layers, 32 / 67 MB sphere maps, a 50 MB cube cross --
    the size_t texel indexing of sample_rgba8 and the 4K frame geometry at the sizes BASELINE names."""
    W, H, spp = 3840, 2160, 2
    sc = S.build_scene(name, W, H, tex_scale=1)
    sizes = [None if t is None else t.shape for t in sc.textures]
    if name == "textures":
        assert sizes[0] == (4, 2048, 2048, 4) and sizes[1][1:] == (2048, 4096, 4)
    if name == "envmap":
        assert sizes[1] == (1, 2048, 4096, 4)
    if name == "cubemap":
        assert sizes[2] == (1, 3072, 4096, 4)
    seeds = S.make_seeds(0x5EED0005, W * H)
    ref, _ = O.trace(sc, seeds, spp, 1, rows=rows)
    for precision in (T.FP32, T.FP64):
        with T.open_scene(sc, spp, seeds, precision=precision) as ctx:
            ctx.trace()
            img = ctx.read().reshape(H, W, 4)
        frac, worst = frac_within(img[rows[0]:rows[1]], ref, TOL[precision])
        assert frac >= 0.999, f"{name} rows {rows} precision {precision}: {frac * 100:.3f}% (worst {worst:.3e})"
        assert np.all(img[..., 3] == 1.0) and not np.isnan(img).any()


def test_config2_and_3_full_frames_are_shard_and_device_invariant():
    """Full 1280x960 frames of configs 2 and 3 at 4 spp: any sharding of the frame reproduces the whole-frame render
    bit for bit (the slices of a pixel are reduced in a fixed order inside the kernel whatever the grid shape)."""
    for name, ap, fl in (("reference", 0.15, 1.6), ("teapot", 0.0, 0.0)):
        W, H, spp = 1280, 960, 4
        sc = S.build_scene(name, W, H, ap, fl)
        seeds = S.make_seeds(0x5EED0002, W * H)
        full = T.render_scene(sc, spp, seeds)
        for shard, count in ((0, 2), (5, 8)):
            with T.open_scene(sc, spp, seeds, shard_index=shard, shard_count=count) as ctx:
                ctx.trace()
                part = ctx.read().reshape(len(ctx.rows), W, 4)
            assert np.array_equal(part, full[ctx.rows]), (name, shard, count)
