"""Fraction of pixels within the BASELINE gates on larger frames (fp32 1e-3, fp64 1e-6), per scene: the margin behind
the 99.9 % gate of the parity tests.   python tools/parity_rate.py [W H spp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from pathtracer_ocl_b200 import scene as S, trace as T
W, H, spp = (int(v) for v in (sys.argv[1:4] + ["320", "240", "2"][len(sys.argv) - 1:]))
for name, ap, fl in (("reference", 0.15, 1.6), ("default", 0, 0), ("transparency", 0, 0), ("textures", 0, 0), ("envmap", 0, 0), ("cubemap", 0, 0),
                     ("teapot", 0, 0), ("gopher", 0, 0), ("christian", 0, 0)):
    sc = S.build_scene(name, W, H, ap, fl, tex_scale=4)
    seeds = S.make_seeds(0xBEEF, W * H)
    ref, _ = O.trace(sc, seeds, spp, precision=1)
    out = []
    for prec, tol in ((T.FP32, 1e-3), (T.FP64, 1e-6)):
        img = T.render_scene(sc, spp, seeds, precision=prec)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        out.append(f"{100 * (err <= tol).mean():8.4f}%")
    print(f"{name:14s} fp32 {out[0]}  fp64 {out[1]}", flush=True)
