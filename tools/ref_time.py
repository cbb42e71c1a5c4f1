"""Kernel-only throughput of the reference scene (headline config) for launch-bounds A/B runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import trace as T
from tools.mesh_time import timing
print(f"reference@512: parity {timing('reference', 1280, 960, 512, 0.15, 1.6, reps=3):8.1f}  transparency@256 {timing('transparency', 1280, 960, 256, reps=2):8.1f}  default@256 {timing('default', 1280, 960, 256, reps=2):8.1f} Mpaths/s", flush=True)
