"""Host phases of one-shot ptc_render calls (PTC_DEBUG_TIMING lines on stderr), for two PTC_BUILD_THREADS settings.
usage: python tools/e2e_phases.py [scene] [W] [H] [spp] [repeats]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PTC_DEBUG_TIMING"] = "1"
import torch  # noqa: E402  (pinned host buffers)
from pathtracer_ocl_b200 import scene as S, trace as T  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "gopher"
W, H, spp, reps = (int(a) for a in (sys.argv[2:6] + ["1280", "960", "2048", "3"][len(sys.argv) - 2:]))
sc = S.build_scene(name, W, H)
seeds = torch.from_numpy(S.make_seeds(1, W * H)).pin_memory()
out = torch.empty(H * W * 4, dtype=torch.float64).pin_memory()
for threads in ("4", "1", "4", "1"):
    os.environ["PTC_BUILD_THREADS"] = threads
    for r in range(reps):
        job = T._Job(sc.objects, sc.triangles if sc.n_triangles else None, sc.groups if sc.n_groups else None, sc.camera,
                     sc.textures[0], sc.textures[1], sc.textures[2], seeds.numpy(), spp, T.FP32, T.RNG_PARITY, [0], 0, 1, 0)
        err = C.create_string_buffer(512)
        t0 = time.perf_counter()
        rc = T.lib().ptc_render(C.byref(job.struct), out.data_ptr(), err, 512)
        print(f"threads {threads} render {r}: rc {rc} wall {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr, flush=True)
