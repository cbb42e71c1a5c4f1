"""Mesh objects on the device: the rebuilt 8-wide BVH and the cooperative walk (trace.cuh: mesh_hit) must pick
exactly the triangle the reference's own walk (tracer.cl:598-742) would pick -- also in the corners where the
reference's box rule differs from geometry: flat node boxes (never passed), direction components below
EPSILON (HUGE_VAL slabs: the whole ancestor chain decides), several group objects sharing one triangle buffer,
and hits at equal distance.  Every case is compared per pixel with the CPU oracle through the C ABI."""
import numpy as np
import pytest

from oracle import oracle as O
from pathtracer_ocl_b200 import scene as S, trace as T

pytestmark = pytest.mark.gpu

TOL = ((T.FP64, 1e-6), (T.FP32, 1e-3))


def set_transform(rec, m):
    inv = np.linalg.inv(m)
    rec["transform"] = m.ravel()
    rec["inverse"] = inv.ravel()
    rec["inverse_transpose"] = inv.T.ravel()


def trs(t=(0, 0, 0), s=(1, 1, 1), ry=0.0):
    m = np.eye(4)
    c, sn = np.cos(ry), np.sin(ry)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, sn, -sn, c
    m[:3, :3] = m[:3, :3] @ np.diag(s)
    m[:3, 3] = t
    return m


def pack(records):
    arr = np.zeros(len(records), dtype=S.OBJECT_DTYPE)
    for i, r in enumerate(records):
        arr[i] = r
    return arr.view(np.uint8).reshape(-1)


def check(sc, seeds, spp, min_tri=1, **kw):
    ref, cnt = O.trace(sc, seeds, spp, precision=1, **({"rng_mode": 1} if kw.get("rng_mode") else {}))
    assert cnt["tri_recorded"] >= min_tri, cnt["tri_recorded"]
    for precision, tol in TOL:
        img = T.render_scene(sc, spp, seeds, precision=precision, **kw)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        frac = float((err <= tol).mean())
        assert frac >= 0.999, f"precision {precision}: {frac * 100:.3f}% of pixels within {tol:g} (worst {err.max():.3e})"
    return ref, cnt


def test_flat_node_boxes_hide_their_triangles_like_upstream():
    """An axis-aligned quad gets a flat BVH node box; tmin < tmax (strict) never holds for it, so the reference never
    tests its triangles.  The tilted quad next to it is visible."""
    W, H = 96, 72
    obj = """
v -1 0 -1
v 1 0 -1
v 1 0 1
v -1 0 1
v -1 0.5 -1
v 1 0.9 -1
v 1 1.3 1
v -1 0.7 1
g flat
f 1 2 3 4
g tilted
f 5 6 7 8
"""
    mesh, _ = S.scene_from_obj(obj, divide_threshold=1)
    room = S.build_scene("default", W, H)
    rv = room.objects_view()
    grp = mesh.objects_view()[0].copy()
    set_transform(grp, trs((0.0, -0.3, 0.0), (0.3, 0.3, 0.3), 0.4))
    grp["color"] = [0.9, 0.3, 0.3, 0.0]
    records = [rv[i] for i in (0, 1, 2, 3, 4, 10)] + [grp]
    sc = S.SceneBuffers("flatbox", W, H, pack(records), mesh.triangles, mesh.groups, room.camera)
    g = sc.groups_view()
    assert ((g["bb_min"][:, :3] == g["bb_max"][:, :3]).any(axis=1)).any(), "the scene is meant to contain a flat node box"
    check(sc, S.make_seeds(11, W * H), 2)


def test_two_meshes_share_one_triangle_buffer_between_analytic_objects():
    """Two group objects (same BVH, different transforms), analytic objects before, between and after them."""
    W, H = 96, 72
    tea = S.build_scene("teapot", W, H)
    tv = tea.objects_view()
    a, b = tv[6].copy(), tv[6].copy()
    set_transform(a, trs((-0.22, -0.4, 0.05), (0.05, 0.05, 0.05), 0.7))
    set_transform(b, trs((0.25, -0.4, -0.1), (0.04, 0.06, 0.04), -1.1))
    b["reflectivity"] = 0.5
    records = [tv[i] for i in range(6)] + [a, tv[7], b]
    sc = S.SceneBuffers("two-teapots", W, H, pack(records), tea.triangles, tea.groups, tea.camera)
    check(sc, S.make_seeds(12, W * H), 2, min_tri=1000)
    check(sc, S.make_seeds(13, W * H), 1, min_tri=1000, rng_mode=T.RNG_FAST)


def test_coincident_meshes_tie_goes_to_the_first_object():
    """The same mesh twice at the same place: every hit distance ties; the reference keeps the first recorded one
    (the lower object index), which decides the material."""
    W, H = 80, 60
    tea = S.build_scene("teapot", W, H)
    tv = tea.objects_view()
    a, b = tv[6].copy(), tv[6].copy()
    a["reflectivity"], b["reflectivity"] = 0.0, 1.0
    for first, second in ((a, b), (b, a)):
        records = [tv[i] for i in range(6)] + [first, second, tv[7]]
        sc = S.SceneBuffers("coincident", W, H, pack(records), tea.triangles, tea.groups, tea.camera)
        check(sc, S.make_seeds(14, W * H), 1, min_tri=1000)


def test_direction_components_below_epsilon_use_the_whole_node_chain():
    """A distant narrow-angle camera looking straight down an axis: in the mesh's object space many primary rays have
    |d.x| or |d.y| < EPSILON, where the reference's slab test degenerates to +-HUGE_VAL and a triangle is reachable
    only if the ray ORIGIN lies inside every ancestor slab.  The device then checks the full chain."""
    W, H = 96, 72
    tea = S.build_scene("teapot", W, H)
    tv = tea.objects_view()
    light = tv[0].copy()                                       # the emissive sphere, blown up to enclose everything
    set_transform(light, trs((0, 0, 0), (600, 600, 600)))
    mesh = tv[6].copy()
    set_transform(mesh, trs((0.2, -1.5, 0.0)))                 # object space == world space up to a shift
    mesh["reflectivity"] = 0.0
    cam = tea.camera.copy()
    cv = cam.view(S.CAMERA_DTYPE)
    half = 0.0075
    cv["half_width"], cv["half_height"] = half, half * H / W
    cv["pixel_size"] = 2 * half / W
    inv = np.eye(4)
    inv[0, 0] = -1.0                                           # looking down +z from z = -500, x mirrored like upstream's view matrix
    inv[2, 2] = -1.0
    inv[2, 3] = -500.0
    cv["inverse"] = inv.ravel()
    sc = S.SceneBuffers("axis-parallel", W, H, pack([light, mesh]), tea.triangles, tea.groups, cam)
    seeds = S.make_seeds(15, W * H)
    ref, cnt = check(sc, seeds, 1, min_tri=1000)
    # the degenerate case really occurs: count primary rays with a sub-EPSILON component
    xs = (np.arange(W) + 0.5) * float(cv["pixel_size"][0]) - half
    assert (np.abs(xs) < 1e-4).sum() >= 1
