"""Where does ptc_render's time go?  Times open / trace / read / close and the one-shot call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pathtracer_ocl_b200 import scene as S, trace as T
W, H = 1280, 960
sc = S.build_scene("reference", W, H, 0.15, 1.6)
seeds = S.make_seeds(1, W * H)
out = np.empty(W * H * 4)
for spp in (1, 1, 64, 2048):
    t0 = time.perf_counter(); ctx = T.open_scene(sc, spp, seeds); t1 = time.perf_counter()
    ctx.trace(); t2 = time.perf_counter(); ctx.read(out); t3 = time.perf_counter(); ctx.close(); t4 = time.perf_counter()
    T.render_scene(sc, spp, seeds); t5 = time.perf_counter()
    print(f"spp {spp}: open {1e3*(t1-t0):.1f} trace {1e3*(t2-t1):.1f} read {1e3*(t3-t2):.1f} close {1e3*(t4-t3):.1f} | one-shot {1e3*(t5-t4):.1f} ms", flush=True)
