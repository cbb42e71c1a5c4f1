"""Histogram of rays per warp that enter a mesh walk (teapot / gopher / cube-map scene at 64 spp), from a tuning build of the
library: tools/ab_build.sh hist "-DPTK_HIST".  The evidence behind the deferral and the private walk."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PTCUDA_LIB"] = os.path.join(ROOT, "build/ab/libptcuda_hist.so")
from pathtracer_ocl_b200 import scene as S, trace as T
import numpy as np
for name in sys.argv[1:] or ["teapot", "gopher", "cubemap"]:
    sc = S.build_scene(name, 1280, 960)
    seeds = S.make_seeds(0x5EED0002, 1280 * 960)
    h = (C.c_ulonglong * 40)()
    T.lib().ptc_debug_hist(None, 1)
    with T.open_scene(sc, 64, seeds) as ctx:
        ctx.trace()
    T.lib().ptc_debug_hist(h, 1)
    a = np.array(h[:33], dtype=np.float64)
    tot = a.sum()
    rounds = sum(a[k] * ((k + 3) // 4) for k in range(33))
    print(name, "calls", int(tot), "P(k=0) %.3f" % (a[0] / tot), "mean k %.2f" % (sum(k * a[k] for k in range(33)) / tot), "rounds/call %.2f" % (rounds / tot))
    print("  share of rounds by k:", " ".join(f"{k}:{100 * a[k] * ((k + 3) // 4) / rounds:.1f}" for k in range(1, 33) if a[k] > 0))
    print("  share of calls  by k:", " ".join(f"{k}:{100 * a[k] / tot:.1f}" for k in range(0, 33) if a[k] > 0))
