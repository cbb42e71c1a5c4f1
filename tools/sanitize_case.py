"""Small renders covering every kernel variant, for compute-sanitizer --tool memcheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T
for name, W, H, spp in [("default", 40, 30, 2), ("teapot", 40, 30, 1), ("gopher", 33, 17, 1), ("textures", 40, 30, 2), ("cubemap", 24, 18, 1), ("reference", 16, 12, 40)]:
    sc = S.build_scene(name, W, H, 0.15 if name == "reference" else 0.0, 1.6 if name == "reference" else 0.0, tex_scale=32)
    seeds = S.make_seeds(7, W * H)
    for prec in (T.FP32, T.FP64):
        for rng in (T.RNG_PARITY, T.RNG_FAST):
            img = T.render_scene(sc, spp, seeds, precision=prec, rng_mode=rng)
    print(name, "ok", float(img[..., :3].mean()))
