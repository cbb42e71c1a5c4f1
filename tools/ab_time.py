"""Time one libptcuda variant (PTCUDA_LIB) on a few workloads; prints one line per workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T

def timing(name, W, H, spp, ap=0.0, fl=0.0, prec=T.FP32, rng=T.RNG_PARITY, reps=3):
    sc = S.build_scene(name, W, H, ap, fl, tex_scale=1)
    seeds = S.make_seeds(0x5EED0002, W * H)
    with T.open_scene(sc, spp, seeds, precision=prec, rng_mode=rng) as ctx:
        best = 1e30
        for _ in range(reps):
            ctx.trace()
            best = min(best, ctx.stats()["kernel_ms"])
        st = ctx.stats()
    return st["paths"] / best / 1e3

tag = os.environ.get("PTCUDA_LIB", "default").split("/")[-1]
res = [f"{timing('reference', 1280, 960, 128, 0.15, 1.6):8.0f}", f"{timing('reference', 1280, 960, 128, 0.15, 1.6, rng=T.RNG_FAST):8.0f}",
       f"{timing('reference', 1280, 960, 32, 0.15, 1.6, prec=T.FP64):8.0f}", f"{timing('teapot', 1280, 960, 8):8.1f}",
       f"{timing('gopher', 1280, 960, 8):8.1f}", f"{timing('transparency', 1280, 960, 64):8.0f}"]
print(f"{tag:32s} ref_fp32 {res[0]}  ref_fast {res[1]}  ref_fp64 {res[2]}  teapot {res[3]}  gopher {res[4]}  transp {res[5]}  (Mpaths/s)", flush=True)
