"""Randomised scenes built directly as wire records (no scene factory): arbitrary affine transforms
(rotations, non-uniform scales, shears), every primitive type and material branch, up to 16 objects.
The CUDA path must agree with the oracle per pixel exactly as on the reference's own scenes."""
import numpy as np
import pytest

from oracle import oracle as O
from pathtracer_ocl_b200 import scene as S, trace as T

pytestmark = pytest.mark.gpu


def rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    i, j = [(1, 2), (0, 2), (0, 1)][axis]
    m[i, i], m[i, j], m[j, i], m[j, j] = c, -s, s, c
    return m


def affine(rng, center, scale, general):
    t = np.eye(4)
    t[:3, 3] = center
    s = np.diag(list(scale) + [1.0])
    if not general:
        return t @ s
    r = rot(0, rng.uniform(0, 6.28)) @ rot(1, rng.uniform(0, 6.28)) @ rot(2, rng.uniform(0, 6.28))
    shear = np.eye(4)
    shear[0, 1] = rng.uniform(-0.3, 0.3)
    return t @ r @ shear @ s


def make_object(kind, m, color, emission=(0, 0, 0), ri=1.0, refl=0.0, min_y=0.0, max_y=0.0):
    rec = np.zeros(1, dtype=S.OBJECT_DTYPE)
    inv = np.linalg.inv(m)
    rec["transform"][0] = m.ravel()
    rec["inverse"][0] = inv.ravel()
    rec["inverse_transpose"][0] = inv.T.ravel()
    rec["color"][0] = list(color) + [0.0]
    rec["emission"][0] = list(emission) + [0.0]
    rec["refractive_index"][0] = ri
    rec["type"][0] = kind
    rec["min_y"][0], rec["max_y"][0] = min_y, max_y
    rec["reflectivity"][0] = refl
    rec["children"][0] = -1
    return rec


def random_scene(seed, W, H):
    rng = np.random.default_rng(seed)
    base = S.build_scene("default", W, H)
    objs = []
    # a closed room so paths bounce: six planes, some of them tilted a little
    for axis, off in ((1, -0.45), (1, 0.45), (0, -0.65), (0, 0.65), (2, 0.5), (2, -2.0)):
        m = np.eye(4)
        m[:3, 3] = [off if a == axis else 0.0 for a in range(3)]
        if axis == 0:
            m = m @ rot(2, np.pi / 2 + rng.uniform(-0.05, 0.05))
        elif axis == 2:
            m = m @ rot(0, np.pi / 2 + rng.uniform(-0.05, 0.05))
        else:
            m = m @ rot(0, rng.uniform(-0.05, 0.05))
        refl = 0.3 if rng.random() < 0.2 else 0.0
        objs.append(make_object(0, m, rng.uniform(0.3, 0.95, 3), refl=refl))
    objs.append(make_object(1, affine(rng, (0, 0.44, -0.2), (0.3, 0.02, 0.3), False), (1, 1, 1), emission=(9, 8, 7)))
    n_extra = int(rng.integers(3, 10))
    for _ in range(n_extra):
        kind = int(rng.integers(1, 4))
        center = (rng.uniform(-0.45, 0.45), rng.uniform(-0.35, 0.2), rng.uniform(-0.6, 0.3))
        scale = rng.uniform(0.05, 0.18, 3)
        material = rng.integers(0, 5)
        kw = {}
        if material == 1:
            kw = dict(refl=float(rng.uniform(0.2, 1.0)))
        elif material == 2:
            kw = dict(ri=float(rng.uniform(1.2, 1.8)), refl=0.05)
        elif material == 3:
            kw = dict(ri=-1.0, refl=0.1)
        if kind == 2:
            kw.update(min_y=0.0, max_y=float(rng.uniform(0.5, 2.0)))
        objs.append(make_object(kind, affine(rng, center, scale, rng.random() < 0.7), rng.uniform(0.2, 1.0, 3), **kw))
    arr = np.zeros(len(objs), dtype=S.OBJECT_DTYPE)      # (np.concatenate would repack the padded dtype)
    for i, o in enumerate(objs):
        arr[i] = o[0]
    objects = arr.view(np.uint8).reshape(-1)
    assert objects.size == 1024 * len(objs)
    return S.SceneBuffers(f"random{seed}", W, H, objects, np.zeros(0, np.uint8), np.zeros(0, np.uint8), base.camera)


@pytest.mark.parametrize("seed", range(12))
def test_random_scene_pixel_parity(seed):
    W, H, spp = 96, 72, 2
    sc = random_scene(seed, W, H)
    assert 10 <= sc.n_objects <= 16
    seeds = S.make_seeds(1000 + seed, W * H)
    ref, cnt = O.trace(sc, seeds, spp, precision=1)
    assert cnt["shaded"] > W * H            # the scene is actually hit
    for precision, tol in ((T.FP64, 1e-6), (T.FP32, 1e-3)):
        img = T.render_scene(sc, spp, seeds, precision=precision)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        frac = float((err <= tol).mean())
        assert frac >= 0.999, f"seed {seed} precision {precision}: {frac * 100:.3f}% within {tol:g} (worst {err.max():.3e})"


def test_bvh_whose_boxes_do_not_contain_their_triangles():
    """The device walk culls by distance only when the host verified that node boxes contain their
    subtrees.  A caller-supplied BVH that violates this must still give the reference's answer: a
    triangle is tested iff the boxes on its node chain are hit, whatever those boxes are."""
    W, H = 96, 72
    sc = S.build_scene("teapot", W, H)
    groups = sc.groups.copy()
    g = groups.view(S.GROUP_DTYPE)
    rng = np.random.default_rng(5)
    for k in rng.choice(len(g), 40, replace=False):          # shrink 40 node boxes towards their centre
        c = 0.5 * (g["bb_min"][k][:3] + g["bb_max"][k][:3])
        g["bb_min"][k][:3] = c + 0.6 * (g["bb_min"][k][:3] - c)
        g["bb_max"][k][:3] = c + 0.6 * (g["bb_max"][k][:3] - c)
    bad = S.SceneBuffers("teapot-shrunk", W, H, sc.objects, sc.triangles, groups, sc.camera)
    seeds = S.make_seeds(77, W * H)
    ref, _ = O.trace(bad, seeds, 1, precision=1)
    good, _ = O.trace(sc, seeds, 1, precision=1)
    assert np.abs(ref - good).max() > 0.1                    # the damage is visible in the oracle's image
    for precision, tol in ((T.FP64, 1e-6), (T.FP32, 1e-3)):
        img = T.render_scene(bad, 1, seeds, precision=precision)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        assert float((err <= tol).mean()) >= 0.999, float((err <= tol).mean())


def test_random_scene_with_mesh_and_fast_rng():
    """A mesh object among random primitives, fast RNG stream, both precisions."""
    W, H, spp = 96, 72, 2
    sc = random_scene(99, W, H)
    teapot = S.build_scene("teapot", W, H)
    grp = teapot.objects.reshape(-1, 1024)[6:7]                     # the teapot group object
    objects = np.concatenate([sc.objects.reshape(-1, 1024)[:12], grp]).reshape(-1)
    sc = S.SceneBuffers("random-mesh", W, H, objects, teapot.triangles, teapot.groups, teapot.camera)
    seeds = S.make_seeds(4242, W * H)
    ref, cnt = O.trace(sc, seeds, spp, precision=1, rng_mode=1)
    assert cnt["tri_recorded"] > 0
    for precision, tol in ((T.FP64, 1e-6), (T.FP32, 1e-3)):
        img = T.render_scene(sc, spp, seeds, precision=precision, rng_mode=T.RNG_FAST)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        assert float((err <= tol).mean()) >= 0.999
