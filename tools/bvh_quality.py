"""CPU-only quality metric of the rebuilt mesh BVH: cooperative-walk steps per ray (inner + leaf iterations of
trace.cuh:mesh_hit) on a ray mix that resembles path tracing -- rays from the room's walls through the mesh's box and
rays leaving the mesh's own surface -- replayed with tests/test_mesh_index.py's model of the device walk."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pathtracer_ocl_b200 import scene as S
import test_mesh_index as M


def rays(m, n, rng):
    row = m["mesh"][m["mesh"][:, 6] >= 0][0]
    lo, hi = row[0:3], row[3:6]
    c, ext = 0.5 * (lo + hi), hi - lo
    tt = m["tri_test"][np.argsort(m["tri_info"][:, 0])]      # in the reference's order, so every build sees the same rays
    out = []
    for i in range(n):
        if i % 2 == 0:                      # from a sphere around the mesh towards a random point of its box
            o = c + 3.0 * np.linalg.norm(ext) * (lambda v: v / np.linalg.norm(v))(rng.normal(size=3))
            d = c + (rng.random(3) - 0.5) * ext - o
        else:                               # leaving the mesh surface in a random direction of the normal's hemisphere
            q = tt[rng.integers(tt.shape[0])]
            p1 = q[0, :3]; e1 = np.array([q[0, 3], q[1, 0], q[1, 1]]); e2 = np.array([q[1, 2], q[1, 3], q[2, 0]])
            u, v = rng.random(2)
            if u + v > 1: u, v = 1 - u, 1 - v
            nrm = np.cross(e1, e2); nrm /= np.linalg.norm(nrm) + 1e-30
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            if d @ nrm < 0: d = -d
            o = p1 + u * e1 + v * e2 + 1e-3 * nrm
        out.append((o, d / np.linalg.norm(d) * 14.0))
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    for name in ("teapot", "gopher"):
        sc = S.build_scene(name, 32, 24)
        m, obj = M.index_of(sc)
        ob = sc.objects_view()[obj]
        obj_lo, obj_hi = np.array(ob["bb_min"][:3]), np.array(ob["bb_max"][:3])
        rng = np.random.default_rng(1)
        stats = {"nodes": 0, "leaves": 0}
        hits = 0
        for o, d in rays(m, n, rng):
            hits += M.replayed_winner(m, m["mesh"][obj], obj_lo, obj_hi, o, d, 1024.0, stats)[0] >= 0
        nodes = m["wide"].shape[0] // 16
        print(f"{name}: {nodes} wide nodes; per ray {stats['nodes'] / n:.2f} inner + {stats['leaves'] / n:.2f} leaf = "
              f"{(stats['nodes'] + stats['leaves']) / n:.2f} steps; {hits}/{n} hit")


if __name__ == "__main__":
    main()
