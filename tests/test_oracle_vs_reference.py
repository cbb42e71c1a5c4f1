"""Pins the oracle on the reference itself.

oracle/tracer_oracle.cpp -- the restatement every GPU parity test compares against -- is checked here against the
reference's OWN kernel source, /root/reference/internal/ocl/tracer.cl, compiled for the host CPU through
oracle/cl_shim.hpp (oracle/build_ref.py; the library lands in oracle/_ref/ and travels to the GPU box, where the
reference source does not exist).  Same scene records, same seeds, same canonical float sin(): the two must agree BIT
FOR BIT on every pixel -- any difference in operation order, a sentinel, a tie-break or a quirk would show up here.
"""
import os
import re

import numpy as np
import pytest

from oracle import build_ref, oracle as O
from pathtracer_ocl_b200 import scene as S

pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="no /root/reference and no prebuilt oracle/_ref/libtracer_ref.so")

CASES = [  # scene, W, H, spp, aperture, focal length
    ("default", 64, 48, 4, 0.0, 0.0),                 # BASELINE config 1 geometry: planes, spheres, cylinder, cube, 0.8 mirror
    ("reference", 64, 48, 8, 0.15, 1.6),              # config 2: depth of field (sunflower lens, NaN ray at sample 0)
    ("teapot", 64, 48, 2, 0.0, 0.0),                  # config 3: BVH walk, interpolated normals
    ("gopher", 48, 36, 1, 0.0, 0.0),                  # config 4: 13 root children, per-triangle colours, unbounded object box
    ("transparency", 64, 48, 4, 0.0, 0.0),            # refraction in / out, Schlick, total internal reflection
    ("transparency_quad_lights", 48, 36, 2, 0.0, 0.0),
    ("transparency_f_light", 48, 36, 2, 0.0, 0.0),
    ("reflection", 48, 36, 4, 0.0, 0.0),
    ("gopher-window", 32, 24, 1, 0.0, 0.0),           # 15 objects
    ("christian", 48, 36, 2, 0.0, 0.0),               # open cylinders
    ("transparent_teapot", 48, 36, 2, 0.0, 0.0),      # thin glass (refractive index -1)
    ("textures", 64, 48, 4, 0.0, 0.0),                # config 5: plane textures, normal maps, sphere maps
    ("envmap", 64, 48, 4, 0.0, 0.0),                  # config 5: emissive textured sky sphere
    ("cubemap", 48, 36, 2, 0.0, 0.0),                 # config 5: cube-cross sky + mesh
]


@pytest.mark.parametrize("name,W,H,spp,ap,fl", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_the_reference_kernel_bit_for_bit(name, W, H, spp, ap, fl):
    sc = S.build_scene(name, W, H, ap, fl, tex_scale=16)
    seeds = S.make_seeds(0x0C1 + W, W * H)
    mine, cnt = O.trace(sc, seeds, spp, precision=1)
    assert cnt["max_intersections"] <= 64, "the reference's 64-slot intersection arrays (tracer.cl:96-102) would overflow"
    ref = O.ref_trace(sc, seeds, spp)
    assert ref[..., :3].max() > 0.0 and np.all(ref[..., 3] == 1.0)            # it really rendered something
    assert np.array_equal(np.isnan(mine), np.isnan(ref))
    same = (mine == ref) | np.isnan(mine)
    assert same.all(), f"{int((~same).any(axis=-1).sum())} of {W * H} pixels differ, worst {np.nanmax(np.abs(mine - ref)):.3e}"


def test_row_ranges_and_thread_counts_do_not_change_the_reference_kernels_output():
    """The driver hands work-items to threads and emulates the 4-scanline batches of ocltracer.go:212-223."""
    W, H, spp = 40, 22, 2                                                     # H is not a multiple of the batch height
    sc = S.build_scene("default", W, H)
    seeds = S.make_seeds(5, W * H)
    full = O.ref_trace(sc, seeds, spp, nthreads=1)
    assert np.array_equal(full, O.ref_trace(sc, seeds, spp, nthreads=7))
    assert np.array_equal(full[6:15], O.ref_trace(sc, seeds, spp, rows=(6, 15)))   # starts and ends inside batches
    mine, _ = O.trace(sc, seeds, spp, precision=1)
    assert np.array_equal(full, mine)


@pytest.mark.skipif(not os.path.exists(build_ref.REFERENCE_KERNEL), reason="reference source not present")
def test_the_only_edit_to_the_reference_source_is_the_vector_literal_spelling():
    src = open(build_ref.REFERENCE_KERNEL, encoding="utf-8").read()
    out = build_ref.translate(src)
    sites = re.findall(r"\((?:double4|double2|float4)\)\s*\(", src)
    assert len(sites) == 38
    # undoing the rewrite gives the source back, character for character
    assert re.sub(r"mk_(double4|double2|float4)\(", lambda m: f"({m.group(1)})(", out) == re.sub(r"\((double4|double2|float4)\)\s*\(", lambda m: f"({m.group(1)})(", src)
    assert "mk_" not in src


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(p for p in os.listdir(GOLDEN) if p.endswith(".npz")))
def test_committed_golden_fixtures_are_the_reference_kernels_output(path):
    """tests/golden/*.npz were written by the oracle (tools/make_golden.py); the reference kernel reproduces every one
    of them bit for bit, so the GPU tests that read them compare against outputs of the reference itself."""
    g = np.load(os.path.join(GOLDEN, path))
    w, h, spp = int(g["width"]), int(g["height"]), int(g["spp"])
    sc = S.build_scene(str(g["scene"]), w, h, float(g["aperture"]), float(g["focal_length"]), tex_scale=int(g["tex_scale"]))
    seeds = S.make_seeds(int(g["seed"]), w * h)
    # the kernel's intersection arrays hold 64 entries and overflow silently (tracer.cl:96-102); on a CPU build that is
    # memory corruption, so the oracle -- which counts them, on the same seeds and samples -- goes first
    mine, cnt = O.trace(sc, seeds, spp, precision=1)
    assert cnt["max_intersections"] <= 64, "fixture scene overflows the reference kernel's fixed arrays"
    assert np.array_equal(mine, g["rgba"])
    ref = O.ref_trace(sc, seeds, spp)
    assert np.array_equal(ref, g["rgba"])


@pytest.mark.parametrize("seed", range(24))
def test_random_scenes_oracle_equals_the_reference_kernel(seed):
    """Randomised scenes built directly as wire records (tests/test_gpu_random_scenes.py: general affine transforms with
    rotations, non-uniform scales and shears; every primitive type and material branch; 10-16 objects): the restatement
    and the reference kernel must still agree bit for bit."""
    from test_gpu_random_scenes import random_scene
    W, H, spp = 48, 36, 2
    sc = random_scene(seed, W, H)
    if seed % 4 == 3:                                  # every fourth scene also gets a mesh object
        teapot = S.build_scene("teapot", W, H)
        grp = teapot.objects.reshape(-1, 1024)[6:7]
        objects = np.concatenate([sc.objects.reshape(-1, 1024)[:12], grp]).reshape(-1)
        sc = S.SceneBuffers("random-mesh", W, H, objects, teapot.triangles, teapot.groups, teapot.camera)
    seeds = S.make_seeds(3000 + seed, W * H)
    mine, cnt = O.trace(sc, seeds, spp, precision=1)
    assert cnt["max_intersections"] <= 64
    ref = O.ref_trace(sc, seeds, spp)
    assert cnt["shaded"] > W * H // 2
    assert np.array_equal(np.isnan(mine), np.isnan(ref))
    same = (mine == ref) | np.isnan(mine)
    assert same.all(), f"seed {seed}: {int((~same).any(axis=-1).sum())} of {W * H} pixels differ, worst {np.nanmax(np.abs(mine - ref)):.3e}"
