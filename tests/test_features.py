"""the two code paths the reference ships switched off -- next-event estimation (tracer.cl:786-825, call
commented out at :1168) and cylinder end caps (:282-310, disabled at :437-444) -- behind ptc_job.features, default off.

Pinned like everything else: oracle/build_ref.py un-comments exactly those call sites of the reference's own source in
memory and compiles the variant for the CPU; the oracle's restatement of the features must equal it BIT FOR BIT (CPU
tests below); the CUDA path must equal the oracle at the BASELINE gates (GPU tests)."""
import os

import numpy as np
import pytest

from oracle import build_ref, oracle as O
from pathtracer_ocl_b200 import scene as S, trace as T

HAVE_SOURCE = os.path.exists(build_ref.REFERENCE_KERNEL)
NEE, CAPS = O.NEE, O.CYLINDER_CAPS
assert (NEE, CAPS) == (T.FEATURE_NEE, T.FEATURE_CYLINDER_CAPS) == (build_ref.NEE, build_ref.CYLINDER_CAPS)

# scene, W, H, spp, features: lights seen through glass and mirrors, cylinders with and without reflectivity, meshes
CASES = [("default", 64, 48, 2, NEE), ("default", 64, 48, 2, CAPS), ("default", 64, 48, 2, NEE | CAPS), ("reference", 64, 48, 3, NEE),
         ("transparency", 64, 48, 2, NEE), ("transparency_quad_lights", 48, 36, 2, NEE), ("christian", 48, 36, 1, NEE | CAPS),
         ("christian", 48, 36, 1, CAPS), ("teapot", 48, 36, 1, NEE), ("textures", 48, 36, 2, NEE)]
IDS = [f"{c[0]}-f{c[4]}" for c in CASES]


@pytest.mark.skipif(not HAVE_SOURCE, reason="needs /root/reference to derive the feature variants")
def test_feature_variants_differ_from_the_stock_source_by_the_comment_markers_only():
    text = open(build_ref.REFERENCE_KERNEL, encoding="utf-8").read()
    stock = text.split("\n")
    for feats, want_changed in ((NEE, 1), (CAPS, 7), (NEE | CAPS, 8)):
        live = build_ref.enable(text, feats).split("\n")
        assert len(live) == len(stock)
        changed = [(a, b) for a, b in zip(stock, live) if a != b]
        assert len(changed) == want_changed
        for a, b in changed:
            assert a.lstrip().startswith("//") and a.replace("//", "", 1).split() == b.split()     # only the marker went away
    assert build_ref.enable(text, 0) == text
    lines = {ln.strip() for ln in build_ref.enable(text, NEE | CAPS).split("\n")}
    assert "nextEventEstimation(objects, numObjects, groups, triangles, &bnce, fgi, fgi2, n, mask, x, &accumColor);" in lines
    assert "double2 caps = intersectCaps(tRayOrigin, tRayDirection, obj.minY, obj.maxY);" in lines


@pytest.mark.parametrize("name,W,H,spp,feats", CASES, ids=IDS)
def test_oracle_features_equal_the_reference_variant_bit_for_bit(name, W, H, spp, feats):
    if O.ref_lib(feats) is None:
        pytest.skip("no compiled reference variant")
    sc = S.build_scene(name, W, H, tex_scale=16)
    seeds = S.make_seeds(0xFEA7 + feats, W * H)
    mine, cnt = O.trace(sc, seeds, spp, precision=1, features=feats)
    if cnt["max_intersections"] > 60:
        pytest.skip("scene overflows the kernel's 64-entry intersection arrays")
    ref = O.ref_trace(sc, seeds, spp, features=feats)
    assert np.array_equal(mine, ref)
    stock, _ = O.trace(sc, seeds, spp, precision=1)
    assert not np.array_equal(mine, stock)                      # the feature is visible in the picture


def test_features_are_off_by_default_and_validated():
    sc = S.build_scene("default", 16, 12)
    job = T._Job(sc.objects, None, None, sc.camera, None, None, None, S.make_seeds(1, 16 * 12), 1, T.FP32, T.RNG_PARITY, None, 0, 1, 0)
    assert job.struct.features == 0
    import ctypes as C
    bad = T._Job(sc.objects, None, None, sc.camera, None, None, None, S.make_seeds(1, 16 * 12), 1, T.FP32, T.RNG_PARITY, None, 0, 1, 0, features=8)
    err = C.create_string_buffer(512)
    h = C.c_void_p()
    assert T.lib().ptc_open(C.byref(bad.struct), C.byref(h), err, 512) != 0 and b"features" in err.value


@pytest.mark.gpu
@pytest.mark.parametrize("name,W,H,spp,feats", CASES + [("gopher", 64, 48, 1, NEE), ("cubemap", 48, 36, 1, NEE)], ids=IDS + ["gopher-f1", "cubemap-f1"])
def test_gpu_features_match_the_oracle(name, W, H, spp, feats):
    sc = S.build_scene(name, W, H, tex_scale=16)
    seeds = S.make_seeds(0xFEA7 + feats, W * H)
    ref, _ = O.trace(sc, seeds, spp, precision=1, features=feats)
    stock = T.render_scene(sc, spp, seeds, precision=T.FP64)
    for precision, tol in ((T.FP64, 1e-6), (T.FP32, 1e-3)):
        img = T.render_scene(sc, spp, seeds, precision=precision, features=feats)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        bad = int((err > tol).sum())
        assert bad <= int(0.001 * W * H) + (1 if precision == T.FP32 else 0), f"precision {precision}: {bad} pixels outside {tol:g} (worst {err.max():.3e})"
        assert np.all(img[..., 3] == 1.0) and not np.isnan(img).any()
    assert not np.array_equal(T.render_scene(sc, spp, seeds, precision=T.FP64, features=feats), stock)


@pytest.mark.gpu
def test_gpu_nee_converges_to_the_oracle_and_needs_the_parity_stream():
    W, H, spp = 48, 36, 256
    sc = S.build_scene("reference", W, H)
    seeds = S.make_seeds(0x7EE, W * H)
    ref, _ = O.trace(sc, seeds, spp, precision=1, features=NEE)
    img = T.render_scene(sc, spp, seeds, precision=T.FP32, features=NEE)
    d = img[..., :3] - ref[..., :3]
    assert np.sqrt(np.mean(d * d)) / np.sqrt(np.mean(ref[..., :3] ** 2)) <= 0.005
    with pytest.raises(T.PtcError, match="PTC_RNG_PARITY"):
        T.render_scene(sc, 1, seeds, rng_mode=T.RNG_FAST, features=NEE)
