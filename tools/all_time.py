"""One line of kernel-only throughputs across the scene families (A/B runs of kernel variants)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import trace as T
from tools.mesh_time import timing
f64 = os.environ.get("ALL_FP64")
if f64:
    print(f"fp64: reference@128 {timing('reference', 1280, 960, 128, 0.15, 1.6, prec=T.FP64):8.1f}  teapot@32 {timing('teapot', 1280, 960, 32, prec=T.FP64):8.1f}  Mpaths/s", flush=True)
else:
    print(f"reference@512 {timing('reference', 1280, 960, 512, 0.15, 1.6):8.1f}  transparency@256 {timing('transparency', 1280, 960, 256):8.1f}  "
          f"teapot@256 {timing('teapot', 1280, 960, 256):8.1f}  gopher@256 {timing('gopher', 1280, 960, 256):8.1f}  textures@128 {timing('textures', 1280, 960, 128):8.1f}  Mpaths/s", flush=True)
