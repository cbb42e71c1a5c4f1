"""The rebuilt mesh BVH (pathtracer_ocl_b200/csrc/ptcuda.cu: BvhBuilder) is only an INDEX: which triangle wins
must still be what the reference's own walk (tracer.cl:598-742) would have picked.  These CPU tests use the
host-only hook `ptc_debug_mesh_index` (no device, no compute) to
  * check the invariants the exactness argument needs from the builder: every triangle in exactly one leaf,
    every stored child box a superset of the triangles below it, depth within the device stack;
  * replay the device walk (trace.cuh: mesh_walk) in numpy, in double, and compare its winner with a
    brute-force statement of the reference rule on the same rays -- including rays with zero direction
    components (|d| < EPSILON -> HUGE_VAL slabs), rays starting inside the mesh and rays grazing flat boxes.
The CUDA code itself is tested against the oracle in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from pathtracer_ocl_b200 import scene as S, trace as T

EPS = 1e-4
WIDE, LEAF_TRIS, WIDE_STACK = 8, 8, 96   # trace.cuh: kWide, kLeafTris, kWideStack
EMPTY = -(1 << 31)                       # trace.cuh: kEmptyChild
SLACK = 1e-13              # trace.cuh: box_slack<double>()


def index_of(scene):
    m = T.debug_mesh_index(scene)
    obj = int(np.flatnonzero(m["mesh"][:, 6] >= 0)[0])
    return m, obj


def leaf_slots(code):
    c = ~int(code)
    return c >> 4, c & 15


def children(m, node):
    """[(code, lo, hi)] of the non-empty children of a wide node."""
    rec = m["wide"][node * WIDE * 2:(node + 1) * WIDE * 2].reshape(WIDE, 2, 4)
    return [(int(rec[c, 0, 3]), rec[c, 0, :3], rec[c, 1, :3]) for c in range(WIDE) if int(rec[c, 0, 3]) != EMPTY]


def subtree(m, child, depth, pending, seen, worst):
    """Returns (lo, hi) of the triangle vertices below `child`; checks stored boxes on the way."""
    worst["depth"] = max(worst["depth"], depth)
    if child < 0:
        first, count = leaf_slots(child)
        assert 1 <= count <= LEAF_TRIS
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        for n in range(first, first + count):
            assert not seen[n]
            seen[n] = True
            q = m["tri_test"][n]
            p1 = q[0, :3]
            e1 = np.array([q[0, 3], q[1, 0], q[1, 1]])
            e2 = np.array([q[1, 2], q[1, 3], q[2, 0]])
            for v in (p1, p1 + e1, p1 + e2):
                lo, hi = np.minimum(lo, v), np.maximum(hi, v)
        return lo, hi
    kids = children(m, child)
    assert 1 <= len(kids) <= WIDE
    worst["stack"] = max(worst["stack"], pending + len(kids) - 1)
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for code, blo, bhi in kids:
        clo, chi = subtree(m, code, depth + 1, pending + len(kids) - 1, seen, worst)
        assert (blo < clo).all() and (bhi > chi).all(), "stored child box must strictly contain its triangles"
        lo, hi = np.minimum(lo, clo), np.maximum(hi, chi)
    return lo, hi


@pytest.mark.parametrize("name", ["teapot", "gopher"])
def test_builder_invariants(name):
    sc = S.build_scene(name, 32, 24)
    m, obj = index_of(sc)
    n = m["tri_info"].shape[0]
    assert n == sc.n_triangles
    assert sorted(m["tri_info"][:, 0].tolist()) == list(range(n)), "ranks are a permutation of the recording order"
    seen = np.zeros(n, dtype=bool)
    worst = {"depth": 0, "stack": 0}
    lo, hi = subtree(m, int(m["mesh"][obj, 6]), 0, 0, seen, worst)
    assert seen.all(), "every triangle sits in exactly one leaf"
    assert worst["stack"] <= WIDE_STACK
    n_nodes = m["wide"].shape[0] // (2 * WIDE)
    print(f"{name}: {n} triangles, {n_nodes} wide nodes, depth {worst['depth']}, worst-case stack {worst['stack']}")
    assert (m["mesh"][obj, 0:3] < lo).all() and (m["mesh"][obj, 3:6] > hi).all()
    # reference nodes: parents precede children (pre-order), triangles point at existing nodes
    par = m["node_parent"]
    assert (par < np.arange(par.size)).all() and (par >= -1).all()
    assert m["tri_info"][:, 1].min() >= 0 and m["tri_info"][:, 1].max() < par.size
    # the reference's recording order: rank increases with the (pre-order) node index
    order = np.argsort(m["tri_info"][:, 0])
    assert (np.diff(m["tri_info"][order, 1]) >= 0).all()


# ---- the reference rule, brute force ---------------------------------------------------------------
def ref_box(o, d, lo, hi):
    """tracer.cl:250-280 for many boxes at once (lo, hi: [n,3]); returns bool[n]."""
    with np.errstate(all="ignore"):
        big = np.abs(d) >= EPS
        a, b = lo - o, hi - o
        t0 = np.where(big, a / np.where(big, d, 1.0), a * np.inf)
        t1 = np.where(big, b / np.where(big, d, 1.0), b * np.inf)
        sw = t0 > t1
        tn, tf = np.where(sw, t1, t0), np.where(sw, t0, t1)
        tmin = np.fmax(np.fmax(tn[:, 0], tn[:, 1]), tn[:, 2])
        tmax = np.fmin(np.fmin(tf[:, 0], tf[:, 1]), tf[:, 2])
        return tmin < tmax


def moller_trumbore(o, d, tt):
    """tracer.cl:640-675 for many triangles; returns (ok, t, u, v)."""
    with np.errstate(all="ignore"):
        p1 = tt[:, 0, :3]
        e1 = np.stack([tt[:, 0, 3], tt[:, 1, 0], tt[:, 1, 1]], axis=1)
        e2 = np.stack([tt[:, 1, 2], tt[:, 1, 3], tt[:, 2, 0]], axis=1)
        dxe2 = np.cross(d[None, :], e2)
        det = e1[:, 0] * dxe2[:, 0] + e1[:, 1] * dxe2[:, 1] + e1[:, 2] * dxe2[:, 2]
        f = 1.0 / det
        sv = o[None, :] - p1
        u = f * (sv[:, 0] * dxe2[:, 0] + sv[:, 1] * dxe2[:, 1] + sv[:, 2] * dxe2[:, 2])
        sxe1 = np.cross(sv, e1)
        v = f * (d[0] * sxe1[:, 0] + d[1] * sxe1[:, 1] + d[2] * sxe1[:, 2])
        t = f * (e2[:, 0] * sxe1[:, 0] + e2[:, 1] * sxe1[:, 1] + e2[:, 2] * sxe1[:, 2])
        ok = ~(np.abs(det) < EPS) & ~((u < 0) | (u > 1)) & ~((v < 0) | ((u + v) > 1))
        return ok, t, u, v


def reference_winner(m, obj_lo, obj_hi, o, d, best_t):
    """Slot of the triangle the reference records as closest with EPS < t < best_t (ties: first recorded), or -1."""
    if not np.isfinite(o).all() or not np.isfinite(d).all():
        return -1, best_t
    if not ref_box(o, d, obj_lo[None, :], obj_hi[None, :])[0]:
        return -1, best_t
    hit = ref_box(o, d, m["node_lo"][:, :3], m["node_hi"][:, :3])
    tested = np.zeros(hit.size, dtype=bool)
    par = m["node_parent"]
    for g in range(hit.size):                       # pre-order: a parent is settled before its children
        tested[g] = hit[g] and (par[g] < 0 or tested[par[g]])
    ok, t, _, _ = moller_trumbore(o, d, m["tri_test"])
    cand = ok & tested[m["tri_info"][:, 1]] & (t > EPS) & (t < best_t)
    if not cand.any():
        return -1, best_t
    idx = np.flatnonzero(cand)
    tmin = t[idx].min()
    tie = idx[t[idx] == tmin]
    return int(tie[np.argmin(m["tri_info"][tie, 0])]), float(tmin)


# ---- the device walk, replayed ------------------------------------------------------------------------
def keep_box(o, k, lo, hi, limit):
    with np.errstate(all="ignore"):
        t0, t1 = (lo - o) * k, (hi - o) * k
        tn = np.fmax(np.fmax(np.fmin(t0[0], t1[0]), np.fmin(t0[1], t1[1])), np.fmax(np.fmin(t0[2], t1[2]), 0.0))
        tf = np.fmin(np.fmin(np.fmax(t0[0], t1[0]), np.fmax(t0[1], t1[1])), np.fmin(np.fmax(t0[2], t1[2]), limit))
        return (not (tn > tf + tf * SLACK)), tn


def chain_ok(m, o, d, g, whole_chain):
    while g >= 0:
        if not ref_box(o, d, m["node_lo"][g:g + 1, :3], m["node_hi"][g:g + 1, :3])[0]:
            return False
        g = int(m["node_parent"][g]) if whole_chain else -1
    return True


def replayed_winner(m, mesh_row, obj_lo, obj_hi, o, d, best_t, stats):
    """trace.cuh mesh_hit for one ray: at an inner node all children are tested, the nearest is entered and the other
    hits are stacked with their entry distance; at a leaf the candidates are tried closest first (rank on ties) and
    the first one the reference would have tested wins; stacked children beyond the best hit are dropped."""
    if not np.isfinite(o).all() or not np.isfinite(d).all():
        return -1, best_t
    if not ref_box(o, d, obj_lo[None, :], obj_hi[None, :])[0]:
        return -1, best_t
    with np.errstate(all="ignore"):
        k = 1.0 / d
    keep, _ = keep_box(o, k, mesh_row[0:3], mesh_row[3:6], best_t * 1.0001)
    if not keep:
        return -1, best_t
    nested = int(mesh_row[7]) & 1
    whole_chain = (not nested) or not (np.abs(d) >= EPS).all()
    ct, crank, cslot = best_t, -1, -1
    cur, stack = int(mesh_row[6]), []
    while cur is not None:
        nxt = None
        if cur >= 0:
            stats["nodes"] += 1
            hits = []
            for lane, (code, lo, hi) in enumerate(children(m, cur)):
                keep, tn = keep_box(o, k, lo, hi, ct * 1.0001)
                if keep:
                    hits.append((tn, lane, code))
            if hits:
                sel = min(hits)                                      # smallest tn, lowest lane on ties
                nxt = sel[2]
                for tn, lane, code in hits:                          # pushed in lane order
                    if lane != sel[1]:
                        stack.append((code, np.float32(np.nextafter(np.float32(tn), np.float32(-np.inf))) if np.float32(tn) > tn else np.float32(tn)))
                assert len(stack) <= WIDE_STACK
        else:
            first, count = leaf_slots(cur)
            stats["leaves"] += 1
            ok, t, _, _ = moller_trumbore(o, d, m["tri_test"][first:first + count])
            cands = []
            for q in range(count):
                n = first + q
                if ok[q] and t[q] > EPS and t[q] <= ct:
                    rank, ref = m["tri_info"][n]
                    if t[q] < ct or rank < crank:
                        cands.append((float(t[q]), int(rank), n, int(ref)))
            for tq, rank, n, ref in sorted(cands):
                if chain_ok(m, o, d, ref, whole_chain):
                    ct, crank, cslot = tq, rank, n
                    break
        if nxt is None:
            while stack:
                code, tn = stack.pop()
                if not (float(tn) > ct * 1.0001):
                    nxt = code
                    break
        cur = nxt
    return cslot, ct


def rays_for(m, obj_lo, obj_hi, rng, n):
    """Object-space rays: random ones aimed at the mesh, axis-parallel ones, ones starting inside, ones aimed at vertices."""
    lo, hi = m["mesh"][:, 0:3][m["mesh"][:, 6] >= 0][0], m["mesh"][:, 3:6][m["mesh"][:, 6] >= 0][0]
    c, ext = 0.5 * (lo + hi), (hi - lo)
    rays = []
    for i in range(n):
        kind = i % 6
        target = c + (rng.random(3) - 0.5) * ext
        if kind == 0:                                   # from outside towards the mesh
            o = c + rng.normal(size=3) * ext * 2.5
            d = target - o
            d /= np.linalg.norm(d) * rng.uniform(0.05, 20.0)          # object-space directions are not unit length
        elif kind == 1:                                 # starting inside the extent, any direction
            o = target
            d = rng.normal(size=3)
        elif kind == 2:                                 # one direction component exactly zero (HUGE_VAL slab)
            o = c + rng.normal(size=3) * ext * 1.5
            d = target - o
            a = rng.integers(3)
            d[a] = 0.0
            o[a] = target[a]
        elif kind == 3:                                 # one component below EPSILON but not zero
            o = c + rng.normal(size=3) * ext * 1.5
            d = target - o
            a = rng.integers(3)
            d[a] = rng.uniform(-0.9e-4, 0.9e-4)
            o[a] = target[a] + rng.uniform(-0.01, 0.01)
        elif kind == 4:                                 # aimed exactly at a triangle vertex (box corners / edges)
            q = m["tri_test"][rng.integers(m["tri_test"].shape[0])]
            o = c + rng.normal(size=3) * ext * 2.0
            d = q[0, :3] - o
        else:                                           # along an axis through the mesh
            a = rng.integers(3)
            o = target.copy()
            o[a] = lo[a] - 1.0
            d = np.zeros(3)
            d[a] = rng.uniform(0.5, 3.0)
        rays.append((o, d))
    return rays


@pytest.mark.parametrize("name,n_rays", [("teapot", 700), ("gopher", 400)])
def test_replayed_walk_picks_the_reference_winner(name, n_rays):
    sc = S.build_scene(name, 32, 24)
    m, obj = index_of(sc)
    ob = sc.objects_view()[obj]
    obj_lo, obj_hi = np.array(ob["bb_min"][:3]), np.array(ob["bb_max"][:3])
    rng = np.random.default_rng(7)
    stats = {"nodes": 0, "leaves": 0}
    hits = 0
    for i, (o, d) in enumerate(rays_for(m, obj_lo, obj_hi, rng, n_rays)):
        best_t = 1024.0 if i % 3 else float(rng.uniform(0.5, 30.0))      # sometimes an analytic hit limits the search
        want = reference_winner(m, obj_lo, obj_hi, o, d, best_t)
        got = replayed_winner(m, m["mesh"][obj], obj_lo, obj_hi, o, d, best_t, stats)
        assert got == want, f"ray {i}: replay {got} vs reference {want} (o={o}, d={d})"
        hits += want[0] >= 0
    assert hits > n_rays // 4
    print(f"{name}: {hits}/{n_rays} rays hit; {stats['nodes'] / n_rays:.1f} inner and {stats['leaves'] / n_rays:.1f} leaf steps per ray "
          f"(the reference walk tests every triangle of every node it enters)")


def flat_box_scene():
    """A mesh whose BVH has a node with a FLAT box (an axis-aligned quad): the reference's strict tmin < tmax never
    passes such a box, so its triangles are invisible upstream -- and must be here."""
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
    return S.scene_from_obj(obj, divide_threshold=1)[0]


def test_flat_reference_boxes_hide_their_triangles():
    sc = flat_box_scene()
    m, obj = index_of(sc)
    ob = sc.objects_view()[obj]
    obj_lo, obj_hi = np.array(ob["bb_min"][:3]), np.array(ob["bb_max"][:3])
    flat = np.flatnonzero((m["node_lo"][:, :3] == m["node_hi"][:, :3]).any(axis=1))
    assert flat.size > 0, "the scene is meant to contain a flat node box"
    rng = np.random.default_rng(3)
    stats = {"nodes": 0, "leaves": 0}
    seen_hidden = 0
    for i in range(300):
        o = np.array([rng.uniform(-0.9, 0.9), rng.uniform(1.5, 3.0), rng.uniform(-0.9, 0.9)])
        tgt = np.array([rng.uniform(-0.9, 0.9), 0.0, rng.uniform(-0.9, 0.9)])
        d = (tgt - o) * rng.uniform(0.2, 2.0)
        want = reference_winner(m, obj_lo, obj_hi, o, d, 1024.0)
        got = replayed_winner(m, m["mesh"][obj], obj_lo, obj_hi, o, d, 1024.0, stats)
        assert got == want
        ok, t, _, _ = moller_trumbore(o, d, m["tri_test"])
        under_flat = np.isin(m["tri_info"][:, 1], flat)
        if (ok & under_flat & (t > EPS)).any():
            seen_hidden += 1
            assert want[0] < 0 or not under_flat[want[0]]
    assert seen_hidden > 100


def geometric_mesh(n, ratio):
    """n small triangles whose positions and sizes shrink geometrically: SAH splits peel them off a few at a time,
    the deepest trees a sane builder produces."""
    lines = []
    for k in range(n):
        x = ratio ** k
        s = 0.3 * x
        lines += [f"v {x} 0 0", f"v {x + s} {s} 0", f"v {x} {s} {s}"]
    lines.append("g spiral")
    lines += [f"f {3 * k + 1} {3 * k + 2} {3 * k + 3}" for k in range(n)]
    return S.scene_from_obj("\n".join(lines) + "\n", divide_threshold=50)[0]


@pytest.mark.parametrize("n,ratio", [(600, 0.95), (3000, 0.99)])
def test_degenerate_mesh_stays_within_the_device_stack(n, ratio):
    sc = geometric_mesh(n, ratio)
    m, obj = index_of(sc)
    seen = np.zeros(m["tri_info"].shape[0], dtype=bool)
    worst = {"depth": 0, "stack": 0}
    subtree(m, int(m["mesh"][obj, 6]), 0, 0, seen, worst)
    assert seen.all()
    assert worst["stack"] <= WIDE_STACK, worst
    print(f"geometric mesh n={n}: depth {worst['depth']}, worst-case stack {worst['stack']}")
    # and the walk still finds the reference's winner
    ob = sc.objects_view()[obj]
    obj_lo, obj_hi = np.array(ob["bb_min"][:3]), np.array(ob["bb_max"][:3])
    rng = np.random.default_rng(5)
    stats = {"nodes": 0, "leaves": 0}
    for i, (o, d) in enumerate(rays_for(m, obj_lo, obj_hi, rng, 120)):
        assert replayed_winner(m, m["mesh"][obj], obj_lo, obj_hi, o, d, 1024.0, stats) == reference_winner(m, obj_lo, obj_hi, o, d, 1024.0)


def test_non_finite_triangles_are_left_out_of_the_index():
    """A triangle with a NaN / infinite coordinate can never be recorded with EPSILON < t < 1024 upstream (every product
    of its test is NaN, inf or 0), so the builder drops it instead of poisoning the boxes; the others stay reachable."""
    sc = S.build_scene("teapot", 32, 24)
    tris = sc.triangles.copy()
    tv = tris.view(S.TRIANGLE_DTYPE)
    bad = [5, 77, 1000, 6319]
    tv["p1"][bad[0]][0] = np.nan
    tv["e1"][bad[1]][1] = np.inf
    tv["e2"][bad[2]][2] = -np.inf
    tv["p1"][bad[3]][1] = np.nan
    broken = S.SceneBuffers("teapot-nan", 32, 24, sc.objects, tris, sc.groups, sc.camera)
    m, obj = index_of(broken)
    assert m["tri_info"].shape[0] == sc.n_triangles - len(bad)
    assert np.isfinite(m["tri_test"]).all() and np.isfinite(m["wide"][:, :3]).all() and np.isfinite(m["mesh"][obj, :6]).all()
    seen = np.zeros(m["tri_info"].shape[0], dtype=bool)
    subtree(m, int(m["mesh"][obj, 6]), 0, 0, seen, {"depth": 0, "stack": 0})
    assert seen.all()
    kept = set(m["tri_info"][:, 0].tolist())                    # ranks keep the reference's numbering, with gaps
    assert len(set(range(sc.n_triangles)) - kept) == len(bad)


def test_two_mesh_objects_get_separate_roots_over_shared_triangles():
    sc = S.build_scene("teapot", 32, 24)
    ov = sc.objects_view()
    recs = np.zeros(3, dtype=S.OBJECT_DTYPE)
    recs[0], recs[1], recs[2] = ov[0], ov[6], ov[6]
    two = S.SceneBuffers("two", 32, 24, recs.view(np.uint8).reshape(-1), sc.triangles, sc.groups, sc.camera)
    m = T.debug_mesh_index(two)
    roots = m["mesh"][:, 6].astype(int)
    assert roots[0] == -1 and roots[1] >= 0 and roots[2] > roots[1]
    n = sc.n_triangles
    assert m["tri_info"].shape[0] == 2 * n                      # each object owns its own slots
    r1, r2 = m["node_range"][1], m["node_range"][2]
    assert r1[1] - r1[0] == r2[1] - r2[0] and r2[0] == r1[1] and r1[1] - r1[0] <= sc.n_groups
    # ranks restart per object; reference nodes of the second object point into its own node range
    assert sorted(m["tri_info"][:n, 0].tolist()) == sorted(m["tri_info"][n:, 0].tolist()) == list(range(n))
    assert m["tri_info"][:n, 1].max() < r1[1] <= m["tri_info"][n:, 1].min()


# ---- the same question put to the reference's own code ---------------------------------------------------------------
from oracle import oracle as _O  # noqa: E402  (test infrastructure)


@pytest.mark.skipif(_O.ref_lib() is None, reason="no compiled reference kernel (oracle/_ref)")
@pytest.mark.parametrize("name,n_rays", [("teapot", 900), ("gopher", 600)])
def test_replayed_walk_picks_the_winner_of_the_reference_kernels_own_walk(name, n_rays):
    """findClosestIntersection of the reference's tracer.cl (compiled for the CPU, oracle/_ref) on a scene that holds
    only the mesh object, untransformed: the replayed device walk must return exactly its t -- also for rays with zero /
    sub-EPSILON direction components, rays from inside the mesh and rays aimed at vertices."""
    sc = S.build_scene(name, 32, 24)
    m, obj = index_of(sc)
    rec = sc.objects_view()[obj:obj + 1].copy()
    eye = np.eye(4).ravel()
    rec["transform"][0], rec["inverse"][0], rec["inverse_transpose"][0] = eye, eye, eye
    only = S.SceneBuffers(name + "-only", 32, 24, rec.view(np.uint8).reshape(-1), sc.triangles, sc.groups, sc.camera)
    obj_lo, obj_hi = np.array(rec["bb_min"][0][:3]), np.array(rec["bb_max"][0][:3])
    rng = np.random.default_rng(21)
    rays = rays_for(m, obj_lo, obj_hi, rng, n_rays)
    ref = _O.ref_closest(only, np.array([np.concatenate([o, d]) for o, d in rays]))
    stats = {"nodes": 0, "leaves": 0}
    hits = 0
    for i, (o, d) in enumerate(rays):
        slot, t = replayed_winner(m, m["mesh"][obj], obj_lo, obj_hi, o, d, 1024.0, stats)
        if ref[i, 1] < 0:
            assert slot < 0, f"ray {i}: the reference hits nothing, the replay hits slot {slot}"
        else:
            hits += 1
            assert slot >= 0 and t == ref[i, 0], f"ray {i}: replay (slot {slot}, t {t!r}) vs reference t {ref[i, 0]!r}"
    assert hits > n_rays // 4


def test_index_does_not_depend_on_the_build_thread_count(monkeypatch):
    """ptc_open builds the halves of large triangle ranges on different threads (BvhBuilder::grow); subtrees are spliced
    back in preorder, so every array of the index must be byte-identical whatever PTC_BUILD_THREADS says."""
    scenes = [S.build_scene("gopher", 32, 24), geometric_mesh(3000, 0.99)]
    for sc in scenes:
        monkeypatch.setenv("PTC_BUILD_THREADS", "1")
        want = T.debug_mesh_index(sc)
        for threads in ("2", "3", "16"):
            monkeypatch.setenv("PTC_BUILD_THREADS", threads)
            got = T.debug_mesh_index(sc)
            assert sorted(got) == sorted(want)
            for k in want:
                assert got[k].tobytes() == want[k].tobytes(), (threads, k)
