"""Kernel-only throughput of the textured scenes (tuning runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import trace as T
from tools.mesh_time import timing
W, H, spp = int(os.environ.get("TEX_W", "1280")), int(os.environ.get("TEX_H", "960")), int(os.environ.get("TEX_SPP", "256"))
print(f"{W}x{H}@{spp}: textures {timing('textures', W, H, spp):8.1f}  envmap {timing('envmap', W, H, spp):8.1f}  cubemap {timing('cubemap', W, H, spp):8.1f}  Mpaths/s", flush=True)
