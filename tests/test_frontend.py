"""Host frontend (libptscene): the reference's own unit-test vectors replayed against the C++
restatement of its Go packages, plus the wire-buffer layout.  CPU only.

Vectors are transcribed from the reference test files cited on each test."""
import ctypes as C
import math

import numpy as np
import pytest

from pathtracer_ocl_b200 import scene as S


def _m(a):
    return (C.c_double * 16)(*a)


def _t(a):
    return (C.c_double * 4)(*a)


def mat_inverse(m):
    out = (C.c_double * 16)()
    S.lib().pts_mat_inverse(_m(m), out)
    return np.array(out[:])


def mat_mul(a, b):
    out = (C.c_double * 16)()
    S.lib().pts_mat_multiply(_m(a), _m(b), out)
    return np.array(out[:])


def transform(kind, x=0.0, y=0.0, z=0.0):
    out = (C.c_double * 16)()
    S.lib().pts_mat_transform(kind.encode(), x, y, z, out)
    return np.array(out[:])


# ---- geom/matrix_test.go:187-254 ------------------------------------------------------------------
INVERSE_CASES = [
    ([-5, 2, 6, -8, 1, -5, 1, 8, 7, 7, -6, -7, 1, -3, 7, 4],
     [0.21805, 0.45113, 0.24060, -0.04511, -0.80827, -1.45677, -0.44361, 0.52068,
      -0.07895, -0.22368, -0.05263, 0.19737, -0.52256, -0.81391, -0.30075, 0.30639]),
    ([8, -5, 9, 2, 7, 5, 6, 1, -6, 0, 9, 6, -3, 0, -9, -4],
     [-0.15385, -0.15385, -0.28205, -0.53846, -0.07692, 0.12308, 0.02564, 0.03077,
      0.35897, 0.35897, 0.43590, 0.92308, -0.69231, -0.69231, -0.76923, -1.92308]),
    ([9, 3, 0, 9, -5, -2, -6, -3, -4, 9, 6, 4, -7, 6, 6, 2],
     [-0.04074, -0.07778, 0.14444, -0.22222, -0.07778, 0.03333, 0.36667, -0.33333,
      -0.02901, -0.14630, -0.10926, 0.12963, 0.17778, 0.06667, -0.26667, 0.33333]),
]


@pytest.mark.parametrize("m,expected", INVERSE_CASES)
def test_inverse_vectors(m, expected):
    got = mat_inverse(m)
    # assert.InEpsilon(expected, actual, geom.Epsilon=0.01): relative error
    assert np.all(np.abs(got - np.array(expected)) <= 0.01 * np.abs(np.array(expected)))


def test_multiply_by_inverse_roundtrip():      # matrix_test.go:238-254
    m1 = [3, -9, 7, 3, 3, -8, 2, -9, -4, 4, 4, 1, -6, 5, -1, 1]
    m2 = [8, 2, 2, 2, 3, -1, 7, 0, 7, 0, 5, 4, 6, -2, 0, 5]
    back = mat_mul(mat_mul(m1, m2), mat_inverse(m2))
    assert np.allclose(back, m1, atol=1e-9)


def test_transform_conventions():              # geom/translation.go:5-12, scaling.go, rotation.go
    t = transform("translate", 5, -3, 2)
    assert (t[3], t[7], t[11]) == (5, -3, 2) and t[15] == 1
    s = transform("scale", 2, 3, 4)
    assert (s[0], s[5], s[10]) == (2, 3, 4)
    r = transform("rotz", math.pi / 2)
    assert r[1] == pytest.approx(-1) and r[4] == pytest.approx(1) and abs(r[0]) < 1e-15
    rx = transform("rotx", math.pi / 2)
    assert rx[6] == pytest.approx(-1) and rx[9] == pytest.approx(1)
    ry = transform("roty", math.pi / 2)
    assert ry[2] == pytest.approx(1) and ry[8] == pytest.approx(-1)


def test_view_transform_default_orientation():  # camera.go:50-81 (Ray Tracer Challenge vectors)
    out = (C.c_double * 16)()
    S.lib().pts_view_transform(_t([0, 0, 0, 1]), _t([0, 0, -1, 1]), _t([0, 1, 0, 0]), out)
    assert np.allclose(out[:], np.eye(4).ravel())
    S.lib().pts_view_transform(_t([0, 0, 8, 1]), _t([0, 0, 0, 1]), _t([0, 1, 0, 0]), out)
    assert np.allclose(out[:], transform("translate", 0, 0, -8))
    S.lib().pts_view_transform(_t([1, 3, 2, 1]), _t([4, -2, 8, 1]), _t([1, 1, 0, 0]), out)
    expect = [-0.50709, 0.50709, 0.67612, -2.36643, 0.76772, 0.60609, 0.12122, -2.82843,
              -0.35857, 0.59761, -0.71714, 0.00000, 0, 0, 0, 1]
    assert np.allclose(out[:], expect, atol=1e-5)


# ---- shapes/bvh_test.go:9-49 ---------------------------------------------------------------------
SPLIT_CASES = [
    ((-1, -4, -5, 9, 6, 5), (-1, -4, -5), (4, 6, 5), (4, -4, -5), (9, 6, 5)),
    ((-1, -2, -3, 9, 5.5, 3), (-1, -2, -3), (4, 5.5, 3), (4, -2, -3), (9, 5.5, 3)),
    ((-1, -2, -3, 5, 8, 3), (-1, -2, -3), (5, 3, 3), (-1, 3, -3), (5, 8, 3)),
    ((-1, -2, -3, 5, 3, 7), (-1, -2, -3), (5, 3, 2), (-1, -2, 2), (5, 3, 7)),
]


@pytest.mark.parametrize("box,lmin,lmax,rmin,rmax", SPLIT_CASES)
def test_split_bounds(box, lmin, lmax, rmin, rmax):
    out = (C.c_double * 16)()
    S.lib().pts_split_bounds(_t(list(box[:3]) + [1]), _t(list(box[3:]) + [1]), out)
    o = np.array(out[:]).reshape(4, 4)[:, :3]
    assert tuple(o[0]) == lmin and tuple(o[1]) == lmax and tuple(o[2]) == rmin and tuple(o[3]) == rmax


# ---- shapes/boundingbox_test.go:203-262 (host slab test, geom.Epsilon = 0.01) ------------------------
BOX1 = ((-1, -1, -1), (1, 1, 1), [
    ((5, 0.5, 0), (-1, 0, 0), True), ((-5, 0.5, 0), (1, 0, 0), True), ((0.5, 5, 0), (0, -1, 0), True),
    ((0.5, -5, 0), (0, 1, 0), True), ((0.5, 0, 5), (0, 0, -1), True), ((0.5, 0, -5), (0, 0, 1), True),
    ((0, 0.5, 0), (0, 0, 1), True), ((-2, 0, 0), (2, 4, 6), False), ((0, -2, 0), (6, 2, 4), False),
    ((0, 0, -2), (4, 6, 2), False), ((2, 0, 2), (0, 0, -1), False), ((0, 2, 2), (0, -1, 0), False),
    ((2, 2, 0), (-1, 0, 0), False)])
BOX2 = ((5, -2, 0), (11, 4, 7), [
    ((15, 1, 2), (-1, 0, 0), True), ((-5, -1, 4), (1, 0, 0), True), ((7, 6, 5), (0, -1, 0), True),
    ((9, -5, 6), (0, 1, 0), True), ((8, 2, 12), (0, 0, -1), True), ((6, 0, -5), (0, 0, 1), True),
    ((8, 1, 3.5), (0, 0, 1), True), ((9, -1, -8), (2, 4, 6), False), ((8, 3, -4), (6, 2, 4), False),
    ((9, -1, -2), (4, 6, 2), False), ((4, 0, 9), (0, 0, -1), False), ((8, 6, -1), (0, -1, 0), False),
    ((12, 5, 4), (-1, 0, 0), False)])


def ray_box_cases():
    for lo, hi, cases in (BOX1, BOX2):
        for o, d, want in cases:
            n = math.sqrt(sum(c * c for c in d))
            yield lo, hi, o, tuple(c / n for c in d), want


@pytest.mark.parametrize("lo,hi,o,d,want", list(ray_box_cases()))
def test_host_ray_box(lo, hi, o, d, want):
    got = S.lib().pts_ray_box(_t(list(o) + [1]), _t(list(d) + [0]), _t(list(lo) + [1]), _t(list(hi) + [1]))
    assert bool(got) == want


# ---- shapes/sphericalmap_test.go:16-23 (host version uses the true pi; exact expectations) -----------
SPHERICAL = [((0, 0, -1), 0.0, 0.5), ((1, 0, 0), 0.25, 0.5), ((0, 0, 1), 0.5, 0.5), ((-1, 0, 0), 0.75, 0.5),
             ((0, 1, 0), 0.5, 1.0), ((0, -1, 0), 0.5, 0.0), ((math.sqrt(2) / 2, math.sqrt(2) / 2, 0), 0.25, 0.75)]


@pytest.mark.parametrize("p,u,v", SPHERICAL)
def test_host_spherical_map(p, u, v):
    uv = (C.c_double * 2)()
    S.lib().pts_spherical_map(_t(list(p) + [1]), uv)
    assert uv[0] == pytest.approx(u, abs=1e-12) and uv[1] == pytest.approx(v, abs=1e-12)


# ---- shapes/cubemap_test.go:9-30 (0 right, 1 left, 2 up, 3 down, 4 front, 5 back) ---------------------
@pytest.mark.parametrize("p,face", [((-1, 0.5, -0.25), 1), ((1.1, -0.75, 0.8), 0), ((0.1, 0.6, 0.9), 4),
                                    ((-0.7, 0, -2), 5), ((0.5, 1, 0.9), 2), ((-0.2, -1.3, 1.1), 3)])
def test_host_cube_face(p, face):
    assert S.lib().pts_cube_face(_t(list(p) + [1])) == face


# ---- obj/objparser_test.go ----------------------------------------------------------------------------
def test_obj_gibberish_is_ignored():           # objparser_test.go:13-21
    text = "There was a young lady named Bright\nwho traveled much faster than light.\nShe set out one day\n" \
           "in a relative way,\nand came back the previous night."
    _, stats = S.scene_from_obj(text)
    assert stats[4] == 5 and stats[3] == 0


def test_obj_vertices_and_faces():             # objparser_test.go:23-58
    text = "\nv -1 1 0\nv -1.0000 0.5000 0.0000\nv 1 0 0\nv 1 1 0\nf 1 2 3\nf 1 3 4\n"
    sb, stats = S.scene_from_obj(text)
    assert stats[0] == 5 and stats[3] == 2      # 4 vertices + the index-0 placeholder
    t = sb.triangles_view()
    assert tuple(t["p1"][0][:3]) == (-1, 1, 0) and tuple(t["p2"][0][:3]) == (-1, 0.5, 0) and tuple(t["p3"][0][:3]) == (1, 0, 0)
    assert tuple(t["p2"][1][:3]) == (1, 0, 0) and tuple(t["p3"][1][:3]) == (1, 1, 0)
    assert np.allclose(t["e1"][0], t["p2"][0] - t["p1"][0]) and np.allclose(t["e2"][0], t["p3"][0] - t["p1"][0])


def test_obj_fan_triangulation():              # objparser_test.go:60-84
    text = "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\nv 0 2 0\nf 1 2 3 4 5"
    sb, stats = S.scene_from_obj(text)
    t = sb.triangles_view()
    assert stats[3] == 3
    verts = [(-1, 1, 0), (-1, 0, 0), (1, 0, 0), (1, 1, 0), (0, 2, 0)]
    for k, (a, b, c) in enumerate([(0, 1, 2), (0, 2, 3), (0, 3, 4)]):
        assert tuple(t["p1"][k][:3]) == verts[a] and tuple(t["p2"][k][:3]) == verts[b] and tuple(t["p3"][k][:3]) == verts[c]


def test_obj_groups():                         # objparser_test.go:86-110
    text = "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\ng FirstGroup\nf 1 2 3\ng SecondGroup\nf 1 3 4"
    sb, stats = S.scene_from_obj(text)
    assert stats[2] == 3                        # DefaultGroup + two named groups
    g = sb.groups_view()
    assert sb.n_groups == 3 and list(g["tri_count"]) == [0, 1, 1]
    assert int(sb.objects_view()["child_count"][0]) == 3


def test_obj_faces_with_normals():             # objparser_test.go:112-150
    text = "v 0 1 0\nv -1 0 0\nv 1 0 0\nvn -1 0 0\nvn 1 0 0\nvn 0 1 0\nf 1//3 2//1 3//2\nf 1/0/3 2/102/1 3/14/2"
    sb, stats = S.scene_from_obj(text)
    t = sb.triangles_view()
    assert stats[1] == 4 and stats[3] == 2
    for k in range(2):
        assert tuple(t["n1"][k][:3]) == (0, 1, 0) and tuple(t["n2"][k][:3]) == (-1, 0, 0) and tuple(t["n3"][k][:3]) == (1, 0, 0)


def test_flat_triangle_normal():               # shapes/triangle.go:21-27: n = normalize(cross(e2, e1))
    sb, _ = S.scene_from_obj("v 0 1 0\nv -1 0 0\nv 1 0 0\nf 1 2 3")
    t = sb.triangles_view()
    assert np.allclose(t["n1"][0][:3], (0, 0, -1)) and np.allclose(t["n2"][0], t["n1"][0]) and np.allclose(t["n3"][0], t["n1"][0])


def test_teapot_model_and_divide():            # objparser_test.go TestProcessModel + SURVEY.md 7.1 counts
    sb = S.build_scene("teapot", 64, 48)
    assert (sb.n_objects, sb.n_triangles, sb.n_groups) == (8, 6320, 269)
    g = sb.groups_view()
    assert int(g["tri_count"].max()) == 305 and int(g["tri_count"].sum()) == 6320
    inner = g["child_group_count"] > 0
    assert int(g["tri_count"][inner].sum()) == 2896           # 46% of the triangles sit in inner nodes
    assert np.all(g["children"][~inner] == 0) and np.all(g["child_group_count"][~inner] == -1)
    # triangles are contiguous per node, in node order (scene.go:113-134)
    assert np.all(np.cumsum(np.r_[0, g["tri_count"][:-1]]) == g["tri_offset"])
    n1 = sb.triangles_view()["n1"][:, :3]
    assert np.allclose(np.linalg.norm(n1, axis=1), 1.0, atol=1e-9)   # computed vertex normals are unit length
    o = sb.objects_view()
    grp = o[o["type"] == 4][0]
    assert int(grp["child_count"]) == 1 and grp["reflectivity"] == 0.2


def test_gopher_model():
    sb = S.build_scene("gopher", 64, 48)
    assert (sb.n_objects, sb.n_triangles, sb.n_groups) == (9, 16640, 516)
    o = sb.objects_view()
    grp = o[o["type"] == 4][0]
    assert int(grp["child_count"]) == 13
    # the empty DefaultGroup widens the root box to +-inf (boundingbox.go:35-60 merge semantics)
    assert np.all(np.isinf(grp["bb_min"][:3])) and np.all(np.isinf(grp["bb_max"][:3]))
    cols = np.unique(np.round(sb.triangles_view()["color"][:, :3], 6), axis=0)
    assert len(cols) >= 5                                       # per-triangle Ka+Kd+Ks colours from gopher.mtl


# ---- scene registry and wire layout ------------------------------------------------------------------
def test_scene_registry_matches_reference_cli():   # cmd/pt/main.go:27-43
    assert S.scene_names() == ["reference", "teapot", "glass", "gopher", "gopher-window", "christian", "textures", "envmap",
                               "cubemap", "reflection", "transparency", "transparency_quad_lights", "transparency_f_light",
                               "transparent_teapot", "default"]


def test_unknown_scene_falls_back_to_default():    # main.go:85-87
    a, b = S.build_scene("no-such-scene", 32, 24), S.build_scene("default", 32, 24)
    assert a.objects.tobytes() == b.objects.tobytes()


def test_glass_scene_reports_missing_asset():       # transparent_glass.go:116 panics on the missing file
    with pytest.raises(RuntimeError, match="glass.obj"):
        S.build_scene("glass", 32, 24)


def test_default_scene_wire_records():              # scenes/ocl.go:114 + scene.go:45-76
    sb = S.build_scene("default", 640, 480)
    o = sb.objects_view()
    assert list(o["type"]) == [0, 0, 0, 0, 0, 1, 1, 2, 3, 4, 1]
    assert int(o["child_count"][9]) == 0            # direct triangle children of a top-level group are dropped
    assert np.all(o["children"] == -1)
    assert (o["min_y"][7], o["max_y"][7]) == (0.0, 0.4)
    assert tuple(o["emission"][10][:3]) == (9, 8, 6) and o["reflectivity"][6] == 0.8
    for k in range(sb.n_objects):                   # inverse really is the inverse; inverse_transpose its transpose
        m, inv = o["transform"][k].reshape(4, 4), o["inverse"][k].reshape(4, 4)
        assert np.allclose(m @ inv, np.eye(4), atol=1e-12)
        assert np.array_equal(o["inverse_transpose"][k].reshape(4, 4), inv.T)
    cam = sb.camera_view()[0]
    assert (int(cam["width"]), int(cam["height"])) == (640, 480)
    half = math.tan(math.pi / 6)
    assert cam["half_width"] == pytest.approx(half) and cam["half_height"] == pytest.approx(half / (640 / 480))
    assert cam["pixel_size"] == pytest.approx(2 * half / 640)


def test_reference_scene_dof_fields():               # scenes/reference.go:16-20, 76
    sb = S.build_scene("reference", 1280, 960, 0.15, 1.6)
    cam = sb.camera_view()[0]
    assert (cam["aperture"], cam["focal_length"]) == (0.15, 1.6)
    assert list(sb.objects_view()["type"]) == [1, 0, 0, 0, 0, 0, 1, 1]


def test_textured_scenes_have_synthetic_textures():
    sb = S.build_scene("textures", 32, 24, tex_scale=16)
    assert sb.textures[0].shape == (4, 128, 128, 4) and sb.textures[1].shape == (2, 128, 256, 4)
    o = sb.objects_view()
    assert int(o["is_textured"].sum()) == 7 and int(o["is_textured_nm"].sum()) == 3
    sb = S.build_scene("cubemap", 32, 24, tex_scale=16)
    assert sb.textures[2].shape == (1, 192, 256, 4) and int(sb.objects_view()["is_env_map"].sum()) == 1


def test_seeds_are_splitmix64_unit_interval():
    s = S.make_seeds(0x5EED0001, 1000)
    assert s.min() >= 0.0 and s.max() < 1.0 and len(np.unique(s)) == 1000
    assert np.array_equal(s, S.make_seeds(0x5EED0001, 1000))
    # first output of splitmix64(seed=0) is 0xE220A8397B1DCDAF
    assert S.make_seeds(0, 1)[0] == (0xE220A8397B1DCDAF >> 11) / 2.0 ** 53


def test_png_and_raw_writers(tmp_path):              # pathtracer.go:32-59, raw/writer.go:11-35
    rgba = np.zeros((2, 3, 4))
    rgba[..., 0] = [[0.0, 0.5, 1.0], [2.0, -1.0, 0.25]]
    rgba[..., 3] = 1
    p = str(tmp_path / "o.png")
    S.write_png(p, rgba, 3, 2)
    from PIL import Image
    im = np.array(Image.open(p))
    assert im.shape == (2, 3, 4) and list(im[0, :, 0]) == [0, 128, 255] and list(im[1, :, 0]) == [255, 0, 64]
    r = str(tmp_path / "o.raw")
    S.write_raw(r, rgba, 3, 2)
    raw = open(r, "rb").read()
    assert len(raw) == 16 + 6 * 12
    assert np.frombuffer(raw[:16], ">i4").tolist() == [1, 0, 3, 2]
    assert np.frombuffer(raw[16:], ">f4").reshape(6, 3)[1, 0] == 0.5
