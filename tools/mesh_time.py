"""Kernel-only throughput of the mesh scenes (and the reference scene as a control) on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T

def timing(name, W, H, spp, ap=0.0, fl=0.0, prec=T.FP32, reps=3):
    sc = S.build_scene(name, W, H, ap, fl, tex_scale=1)
    seeds = S.make_seeds(0x5EED0002, W * H)
    with T.open_scene(sc, spp, seeds, precision=prec) as ctx:
        best = 1e30
        for _ in range(reps):
            ctx.trace()
            best = min(best, ctx.stats()["kernel_ms"])
        st = ctx.stats()
    return st["paths"] / best / 1e3

if __name__ == "__main__":
  spp = int(os.environ.get("MESH_SPP", "32"))
  print(f"teapot {timing('teapot', 1280, 960, spp):8.1f}  gopher {timing('gopher', 1280, 960, spp):8.1f}  "
        f"cubemap {timing('cubemap', 1280, 960, spp):8.1f}  reference {timing('reference', 1280, 960, 128, 0.15, 1.6):8.1f}  Mpaths/s (fp32)", flush=True)
  if os.environ.get("MESH_FP64"):
    print(f"fp64: teapot {timing('teapot', 1280, 960, 8, prec=T.FP64):8.1f}  gopher {timing('gopher', 1280, 960, 8, prec=T.FP64):8.1f}  Mpaths/s", flush=True)
