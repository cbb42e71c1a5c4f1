"""Mesh-kernel tuning sweep on the GPU box: block size x drain threshold, teapot and gopher (Mpaths/s)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T

def timing(sc, seeds, spp, reps=3, prec=T.FP32):
    with T.open_scene(sc, spp, seeds, precision=prec) as ctx:
        best = 1e30
        for _ in range(reps):
            ctx.trace()
            best = min(best, ctx.stats()["kernel_ms"])
        st = ctx.stats()
    return st["paths"] / best / 1e3

W, H = 1280, 960
seeds = S.make_seeds(0x5EED0002, W * H)
scenes = {n: S.build_scene(n, W, H, 0.0, 0.0, tex_scale=1) for n in ("teapot", "gopher")}
spp = int(os.environ.get("SWEEP_SPP", "32"))
blocks = [int(x) for x in os.environ.get("SWEEP_BLOCKS", "128,256").split(",")]
drains = [int(x) for x in os.environ.get("SWEEP_DRAINS", "1,32,64,96,128,192").split(",")]
for block in blocks:
    for drain in drains:
        if drain > block:
            continue
        os.environ["PTC_MESH_BLOCK"], os.environ["PTC_MESH_DRAIN"] = str(block), str(drain)
        res = {n: timing(sc, seeds, spp) for n, sc in scenes.items()}
        print(f"block {block:4d} drain {drain:4d}  teapot {res['teapot']:8.1f}  gopher {res['gopher']:8.1f}  Mpaths/s", flush=True)
if os.environ.get("SWEEP_FP64"):
    for n, sc in scenes.items():
        print(f"fp64 {n} {timing(sc, seeds, 8, prec=T.FP64):8.1f} Mpaths/s", flush=True)
