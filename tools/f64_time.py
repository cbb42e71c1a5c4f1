"""fp64-mode throughput (reference scene, teapot) for launch-bounds A/B runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T
from tools.mesh_time import timing

print(f"fp64: reference {timing('reference', 1280, 960, 64, 0.15, 1.6, prec=T.FP64):8.1f}  teapot {timing('teapot', 1280, 960, 16, prec=T.FP64):8.1f}  Mpaths/s", flush=True)
