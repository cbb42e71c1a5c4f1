"""Kernel-only throughput of the BASELINE configs at their full sample counts (CUDA-event time, best of `reps`).

    python tools/r2_time.py [ref] [teapot] [gopher] [ref64] [teapot64] [tex] [env] [cube] ...   (default: ref teapot gopher)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T

CASES = {
    "ref": ("reference", 1280, 960, 2048, 0.15, 1.6, T.FP32), "ref64": ("reference", 1280, 960, 512, 0.15, 1.6, T.FP64),
    "teapot": ("teapot", 1280, 960, 2048, 0, 0, T.FP32), "teapot64": ("teapot", 1280, 960, 256, 0, 0, T.FP64),
    "gopher": ("gopher", 1280, 960, 2048, 0, 0, T.FP32), "gopher64": ("gopher", 1280, 960, 256, 0, 0, T.FP64),
    "tex": ("textures", 1280, 960, 512, 0, 0, T.FP32), "env": ("envmap", 1280, 960, 512, 0, 0, T.FP32),
    "cube": ("cubemap", 1280, 960, 512, 0, 0, T.FP32), "transp": ("transparency", 1280, 960, 512, 0, 0, T.FP32),
    "default": ("default", 1280, 960, 512, 0, 0, T.FP32),
}

def timing(name, W, H, spp, ap, fl, prec, reps=2):
    sc = S.build_scene(name, W, H, ap, fl, tex_scale=1)
    seeds = S.make_seeds(0x5EED0002, W * H)
    with T.open_scene(sc, spp, seeds, precision=prec) as ctx:
        best = 1e30
        for _ in range(reps):
            ctx.trace()
            best = min(best, ctx.stats()["kernel_ms"])
        st = ctx.stats()
    return st["paths"] / best / 1e3, best

if __name__ == "__main__":
    names = sys.argv[1:] or ["ref", "teapot", "gopher"]
    spp_scale = float(os.environ.get("SPP_SCALE", "1"))
    out = []
    for n in names:
        c = list(CASES[n]); c[3] = max(1, int(c[3] * spp_scale))
        mp, ms = timing(*c)
        out.append(f"{n} {mp:8.1f} ({ms:7.1f} ms)")
    print(T.lib().ptc_version().decode(), "|", "  ".join(out), " Mpaths/s", flush=True)
